/* placeholder until the warp-specialised kernel lands */
#include "kernels.h"
namespace lrpt {
bool ws_supported(const lrpt_consts_t &) { return false; }
cudaError_t ws_prepare(int) { return cudaSuccess; }
cudaError_t launch_ws(const LaunchArgs &, cudaStream_t, int *n) { if (n) *n = 0; return cudaErrorNotSupported; }
}
