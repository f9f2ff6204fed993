// Mailbox round-trip variants between two warps on different SM sub-partitions (cycles per round trip).
#include <cstdio>
#include <cuda_runtime.h>
#define ROUNDS 4096
__device__ __forceinline__ unsigned sm(const volatile void *p) { return (unsigned)__cvta_generic_to_shared((const void *)p); }
__device__ __forceinline__ int ldv(const volatile int *p) { int v; asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(sm(p)) : "memory"); return v; }
__device__ __forceinline__ void stv(volatile int *p, int v) { asm volatile("st.volatile.shared.b32 [%0], %1;" :: "r"(sm(p)), "r"(v) : "memory"); }

template <int MODE> __device__ __forceinline__ void wait_for(const volatile int *slot, int want)
{
	if (MODE == 0) { while (__any_sync(0xffffffffu, ldv(slot) != want)) { } }              // vote loop (current)
	else if (MODE == 1) { while (ldv(slot) != want) { } __syncwarp(); }                      // per-lane exit + reconverge
	else if (MODE == 2) {                                                                    // two loads in flight per check
		int a = ldv(slot);
		for (;;) { const int b = ldv(slot); if (a == want) break; a = b; }
		__syncwarp();
	} else if (MODE == 3) {                                                                  // four loads per branch, no vote
		for (;;) {
			const int a = ldv(slot), b = ldv(slot), c = ldv(slot), d = ldv(slot);
			if (a == want || b == want || c == want || d == want) break;
		}
		__syncwarp();
	}
}
template <int MODE> __global__ void k(long long *out)
{
	__shared__ volatile int a[32], b[32];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x < 32) { a[lane] = 0; b[lane] = 0; }
	__syncthreads();
	const long long t0 = clock64();
	if (warp == 0) { for (int i = 1; i <= ROUNDS; i++) { stv(&a[lane], i); wait_for<MODE>(&b[lane], i); } }
	else if (warp == 1) { for (int i = 1; i <= ROUNDS; i++) { wait_for<MODE>(&a[lane], i); stv(&b[lane], i); } }
	const long long t1 = clock64();
	if (threadIdx.x == 0) out[0] = t1 - t0;
}
// mbarrier-based hand-over: writer arrives, reader try_waits on the phase
__global__ void k_mbar(long long *out)
{
	__shared__ unsigned long long ba, bb;
	__shared__ volatile int a[32], b[32];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" :: "r"(sm(&ba)));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" :: "r"(sm(&bb)));
	}
	__syncthreads();
	const long long t0 = clock64();
	for (int i = 1; i <= ROUNDS; i++) {
		const unsigned par = (unsigned)(i - 1) & 1u;
		if (warp == 0) {
			a[lane] = i;
			asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" :: "r"(sm(&ba)) : "memory");
			asm volatile("{\n.reg .pred p;\nW0_%=: mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n@!p bra W0_%=;\n}" :: "r"(sm(&bb)), "r"(par) : "memory");
		} else if (warp == 1) {
			asm volatile("{\n.reg .pred p;\nW1_%=: mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n@!p bra W1_%=;\n}" :: "r"(sm(&ba)), "r"(par) : "memory");
			b[lane] = a[lane];
			asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" :: "r"(sm(&bb)) : "memory");
		}
	}
	const long long t1 = clock64();
	if (threadIdx.x == 0) out[0] = t1 - t0;
}
int main()
{
	long long *d, h;
	cudaMalloc(&d, 8);
#define RUN(K, name) do { K<<<1, 64>>>(d); cudaDeviceSynchronize(); K<<<1, 64>>>(d); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); \
	printf("%-34s %7.1f cycles per round trip%s\n", name, (double)h/ROUNDS, cudaGetLastError() ? " (error)" : ""); } while (0)
	RUN(k<0>, "vote loop (round-2 mailbox)"); RUN(k<1>, "per-lane exit + syncwarp"); RUN(k<2>, "two loads in flight");
	RUN(k<3>, "four loads per branch"); RUN(k_mbar, "mbarrier arrive / try_wait");
	return 0;
}
