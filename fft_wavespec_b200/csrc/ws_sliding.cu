// ws_sliding.cu — shared-butterfly sliding FFT (placeholder until the kernel lands).
#include "ws_common.cuh"
#include "ws_series.h"
namespace ws {
bool sliding_shared_supported(const Params&) { return false; }
cudaError_t launch_sliding_shared(Params, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace ws
