// agc_scan.cuh -- label utilities (min/max, unpack) and the N->N family (scans, sort).
#pragma once
#include "agc_generic.cuh"
namespace agc {

__global__ void k_minmax_init(long long* mm) { mm[0] = 0x7fffffffffffffffll; mm[1] = (long long)0x8000000000000000ll; }

__global__ void k_label_minmax(const void* labels, int dtype, long long n, long long* mm) {
  long long lo = 0x7fffffffffffffffll, hi = (long long)0x8000000000000000ll;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    long long v = load_int(labels, dtype, i);
    if (dtype == AGC_DT_U64 && v < 0) v = 0x7fffffffffffffffll;     // >= 2^63: saturate, the caller rejects it
    lo = v < lo ? v : lo; hi = v > hi ? v : hi;
  }
  for (int o = 16; o; o >>= 1) {
    long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(mm, lo); atomicMax(mm + 1, hi); }
}

__global__ void k_unpack(const void* table, int eb, long long size, const void* labels, int dtype, long long n,
                         void* out, int* status) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  bool bad = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    long long g = load_int(labels, dtype, i);
    if (g < 0 || g >= size) { bad = true; continue; }
    switch (eb) {
      case 1: ((unsigned char*)out)[i] = ((const unsigned char*)table)[g]; break;
      case 2: ((unsigned short*)out)[i] = ((const unsigned short*)table)[g]; break;
      case 4: ((unsigned*)out)[i] = ((const unsigned*)table)[g]; break;
      default: ((unsigned long long*)out)[i] = ((const unsigned long long*)table)[g]; break;
    }
  }
  if (bad) status[AGC_ST_BADLABEL] = 1;
}

inline const char* scan_error() { return "N->N operations are not built yet"; }
inline int64_t scan_workspace_bytes(int, int64_t, int64_t, int) { return 0; }
inline int64_t sort_workspace_bytes(int64_t, int64_t) { return 0; }
inline int run_scan_op(const agc_request_t*, cudaStream_t, int) { return -20; }
inline int group_order(const agc_index_t&, int64_t, int64_t, long long*, long long*, void*, int*, cudaStream_t, int) { return -20; }

}  // namespace agc
