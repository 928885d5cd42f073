// agc_fast.cuh -- specialised shared-memory kernels (placeholder until the first GPU parity run).
#pragma once
#include "agc_generic.cuh"
namespace agc {
// returns 1 if it handled accumulation into the global tables, 0 to fall back, <0 on error
inline int try_fast_path(const agc_request_t*, const Tables&, cudaStream_t, int) { return 0; }
}  // namespace agc
