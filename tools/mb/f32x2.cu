// Microbenchmark behind demod_lane.cu's packed tap arithmetic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2.bin f32x2.cu && ./f32x2.bin
// Measured on a B200: (1) packed FMUL2/FFMA2/FADD2 issue at half the rate of scalar fp32 (same flops per
// cycle, half the issue slots); (2) ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with
// --fmad=false, and also folds fma(x,h,-0)+add and mul+fma(acc,1.0,p) into it (modes 1-3 differ from the
// scalar chain in 31 % of the results); (3) mode 4, mul2 + fma2(p, ONE, acc) with ONE = 1.0f passed as a
// kernel argument, is bit-identical to the scalar chain.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <stdint.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 r; asm("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}

// MODE 4: mul2 + fma2(p, ONE_runtime, acc)
// MODE 0: scalar mul+add ; 1: mul2+add2 (ptxas may fuse) ; 2: fma2(x,h,-0)+add2 ; 3: mul2 + fma2(acc,1,p)
template<int MODE>
__global__ void k(float *out, const float* in, int iters, float one)
{
    float x0=in[threadIdx.x], x1=in[threadIdx.x+32], h=in[64+ (threadIdx.x&7)];
    float a[8]; for(int i=0;i<8;i++) a[i]=in[i+threadIdx.x];
    if (MODE==0) {
        for (int it=0; it<iters; it++) {
            #pragma unroll
            for (int j=0;j<4;j++){ a[2*j]=__fadd_rn(a[2*j], __fmul_rn(x0,h)); a[2*j+1]=__fadd_rn(a[2*j+1], __fmul_rn(x1,h)); x0=__fmul_rn(x0,1.0000001f); x1=__fmul_rn(x1,0.9999999f);}
        }
    } else {
        u64 A[4]; for(int j=0;j<4;j++) A[j]=pk(a[2*j],a[2*j+1]);
        u64 X=pk(x0,x1), H=pk(h,h), G=pk(1.0000001f,0.9999999f), NZ=pk(-0.0f,-0.0f), ONE=pk(1.0f,1.0f), ONER=pk(one,one);
        for (int it=0; it<iters; it++) {
            #pragma unroll
            for (int j=0;j<4;j++){
                if (MODE==1) A[j]=add2(A[j], mul2(X,H));
                if (MODE==2) A[j]=add2(A[j], fma2(X,H,NZ));
                if (MODE==3) A[j]=fma2(A[j], ONE, mul2(X,H));
                if (MODE==4) A[j]=fma2(mul2(X,H), ONER, A[j]);
                X=mul2(X,G);}
        }
        for(int j=0;j<4;j++) upk(A[j],a[2*j],a[2*j+1]);
    }
    for(int i=0;i<8;i++) out[(blockIdx.x*blockDim.x+threadIdx.x)*8+i]=a[i];
}
int main(){
    const int NT=148*1024;
    float *in,*out[5]; cudaMalloc(&in,4096); for(int m=0;m<5;m++) cudaMalloc(&out[m],NT*8*4);
    float hin[1024]; for(int i=0;i<1024;i++) hin[i]=0.37f+0.013f*i; cudaMemcpy(in,hin,4096,cudaMemcpyHostToDevice);
    int iters=20000;
    for (int warps=4; warps<=32; warps*=2) for(int mode=0; mode<5; mode++){
        cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float ms=0;
        for(int rep=0;rep<2;rep++){
        cudaEventRecord(e0);
        if(mode==0) k<0><<<148,warps*32>>>(out[0],in,iters,1.0f); else if(mode==1) k<1><<<148,warps*32>>>(out[1],in,iters,1.0f); else if(mode==4) k<4><<<148,warps*32>>>(out[4],in,iters,1.0f);
        else if(mode==2) k<2><<<148,warps*32>>>(out[2],in,iters,1.0f); else k<3><<<148,warps*32>>>(out[3],in,iters,1.0f);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);}
        double lane_ops = 148.0*warps*32*iters*4*6;
        printf("mode %d warps %d: %.3f ms, %.2f T lane-ops/s (%s)\n",mode,warps,ms,lane_ops/ms/1e9,cudaGetErrorString(cudaGetLastError()));
    }
    static float h[5][148*1024*8];
    for(int m=0;m<5;m++) cudaMemcpy(h[m],out[m],NT*8*4,cudaMemcpyDeviceToHost);
    for(int m=1;m<5;m++){ long diff=0; for(int i=0;i<NT*8;i++) diff += (memcmp(&h[m][i],&h[0][i],4)!=0); printf("mode %d vs scalar: %ld of %d values differ\n",m,diff,NT*8);}
}
