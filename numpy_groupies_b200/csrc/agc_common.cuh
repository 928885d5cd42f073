// agc_common.cuh -- shared device helpers for libaggcuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <atomic>
#include "../../include/aggcuda.h"

#define AGC_EMPTY_MAX  ((long long)0x8000000000000000ll)   // "no element yet" for max-keys
#define AGC_EMPTY_MIN  ((long long)0x7fffffffffffffffll)   // "no element yet" for min-keys
#define AGC_NO_INDEX   ((long long)0x7fffffffffffffffll)

// status word slots
#define AGC_ST_BADLABEL 0
#define AGC_ST_PRECISION 1
#define AGC_ST_UNSORTED 2
#define AGC_ST_REGIME 3

namespace agc {

// kernels launched by this library so far (bench.py reports it as `gpu_launches`)
static std::atomic<long long> g_launches{0};
#define AGC_COUNT_LAUNCH(n_) (::agc::g_launches.fetch_add((n_), std::memory_order_relaxed))

// ---- per-DEVICE host state (a process may drive several GPUs, from several threads) ----------------------------------
constexpr int AGC_MAX_DEVICES = 64;
inline int current_device() { int dev = 0; cudaGetDevice(&dev); return (dev < 0 || dev >= AGC_MAX_DEVICES) ? 0 : dev; }
// SM count of the CURRENT device; grid cap for the tile-loop kernels
inline int device_sms() {
  static std::atomic<int> sms[AGC_MAX_DEVICES];
  const int dev = current_device();
  int v = sms[dev].load(std::memory_order_relaxed);
  if (!v) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); if (v <= 0) v = 148; sms[dev].store(v, std::memory_order_relaxed); }
  return v;
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device / context: `done` remembers the devices it was set on.
// Returns false if the attribute cannot be set.
template <class K>
inline bool ensure_dyn_smem(K kern, int bytes, std::atomic<unsigned long long>& done) {
  const int dev = current_device();
  if (done.load(std::memory_order_acquire) >> dev & 1ull) return true;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
  done.fetch_or(1ull << dev, std::memory_order_release);
  return true;
}
inline int tile_grid(long long ntiles, int ctas_per_sm) {
  const long long cap = (long long)device_sms() * ctas_per_sm;
  return (int)(ntiles < 1 ? 1 : (ntiles < cap ? ntiles : cap));
}

// ---------------------------------------------------------------------------------------------
// value classes: every input dtype is widened at load time to one of four classes
// ---------------------------------------------------------------------------------------------
struct ClsF32 { using type = float;  static constexpr bool is_float = true;  };
struct ClsF64 { using type = double; static constexpr bool is_float = true;  };
struct ClsI64 { using type = long long; static constexpr bool is_float = false; };
struct ClsU64 { using type = unsigned long long; static constexpr bool is_float = false; };

__host__ __device__ inline int dtype_bytes(int dt) {
  switch (dt) {
    case AGC_DT_BOOL: case AGC_DT_I8: case AGC_DT_U8: return 1;
    case AGC_DT_I16: case AGC_DT_U16: return 2;
    case AGC_DT_I32: case AGC_DT_U32: case AGC_DT_F32: return 4;
    default: return 8;
  }
}
__host__ __device__ inline bool dtype_is_float(int dt) { return dt == AGC_DT_F32 || dt == AGC_DT_F64; }

// integer load with runtime dtype (warp-uniform switch)
__device__ __forceinline__ long long load_int(const void* p, int dt, long long i) {
  switch (dt) {
    case AGC_DT_BOOL: return ((const unsigned char*)p)[i] != 0;
    case AGC_DT_I8:  return ((const signed char*)p)[i];
    case AGC_DT_U8:  return ((const unsigned char*)p)[i];
    case AGC_DT_I16: return ((const short*)p)[i];
    case AGC_DT_U16: return ((const unsigned short*)p)[i];
    case AGC_DT_I32: return ((const int*)p)[i];
    case AGC_DT_U32: return ((const unsigned*)p)[i];
    default:         return ((const long long*)p)[i];      // I64 and U64 (bit pattern)
  }
}
// wrap an int64 to the value range of integer dtype dt (numpy's same-width arithmetic)
__device__ __forceinline__ long long wrap_int(long long v, int dt) {
  switch (dt) {
    case AGC_DT_BOOL: return v != 0;
    case AGC_DT_I8:  return (signed char)v;
    case AGC_DT_U8:  return (unsigned char)v;
    case AGC_DT_I16: return (short)v;
    case AGC_DT_U16: return (unsigned short)v;
    case AGC_DT_I32: return (int)v;
    case AGC_DT_U32: return (unsigned)v;
    default: return v;
  }
}

// ---------------------------------------------------------------------------------------------
// order-preserving signed keys (min/max/arg*): signed integer compare == value compare
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long key_of(double v) {       // finite, +-inf; NaN handled by the caller
  long long b = __double_as_longlong(v);
  return b ^ ((b >> 63) & 0x7fffffffffffffffll);
}
__device__ __forceinline__ double unkey_f64(long long k) {
  return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffll));
}
__device__ __forceinline__ long long key_of(float v) { return key_of((double)v); }
__device__ __forceinline__ long long key_of(long long v) { return v; }
__device__ __forceinline__ long long key_of(unsigned long long v) { return (long long)(v ^ 0x8000000000000000ull); }
__device__ __forceinline__ int key32_of(float v) { int b = __float_as_int(v); return b ^ ((b >> 31) & 0x7fffffff); }
__device__ __forceinline__ float unkey_f32(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

template <class T> __device__ __forceinline__ bool is_nan(T) { return false; }
template <> __device__ __forceinline__ bool is_nan<float>(float v) { return v != v; }
template <> __device__ __forceinline__ bool is_nan<double>(double v) { return v != v; }

// ---------------------------------------------------------------------------------------------
// flat group of element i (Forms 1-5 folded into the index-load stage)
// ---------------------------------------------------------------------------------------------
struct Indexer {
  agc_index_t ix;
  long long size;        // flat number of groups
  long long n;           // number of elements (MULTI: row pitch of the label matrix)
  // returns the flat group, or -1 for a bad label / coordinate
  __device__ __forceinline__ long long group(long long i) const {
    if (ix.form == AGC_FORM_FLAT) {
      long long g = load_int(ix.labels, ix.dtype, i);
      return (g < 0 || g >= size) ? -1 : g;
    } else if (ix.form == AGC_FORM_AXIS) {
      long long rem = i, g = 0;
      for (int d = ix.ndim - 1; d >= 0; --d) {
        long long dim = ix.dims[d];
        long long c = rem % dim; rem /= dim;
        if (d == ix.axis) {
          c = load_int(ix.labels, ix.dtype, c);
          if (c < 0 || c >= ix.axis_size) return -1;
        }
        g += c * ix.strides[d];
      }
      return g;
    } else {
      long long g = 0;
      for (int d = 0; d < ix.ndim; ++d) {
        long long c = load_int(ix.labels, ix.dtype, (long long)d * n + i);
        if (c < 0 || c >= ix.dims[d]) return -1;     // also rejects uint64 labels >= 2^63
        g += c * ix.strides[d];
      }
      return g;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// numpy-compatible casts for the finalise stage
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long f64_to_i64_x86(double x) {   // cvttsd2si: NaN / out of range -> INT64_MIN
  if (!(x > -9.2233720368547758e18 && x < 9.2233720368547758e18)) return (long long)0x8000000000000000ll;
  return (long long)x;
}
__device__ __forceinline__ void store_from_f64(void* out, int dt, long long i, double x) {
  switch (dt) {
    case AGC_DT_F64: ((double*)out)[i] = x; break;
    case AGC_DT_F32: ((float*)out)[i] = (float)x; break;
    case AGC_DT_BOOL: ((unsigned char*)out)[i] = (x != 0.0); break;        // NaN is truthy
    case AGC_DT_I64: ((long long*)out)[i] = f64_to_i64_x86(x); break;
    case AGC_DT_U64: ((unsigned long long*)out)[i] = x >= 9.2233720368547758e18 ? (unsigned long long)x
                                                      : (unsigned long long)f64_to_i64_x86(x); break;
    case AGC_DT_I32: {
      int v = (!(x > -2147483649.0 && x < 2147483648.0)) ? (int)0x80000000 : (int)x;
      ((int*)out)[i] = v; break; }
    case AGC_DT_U32: ((unsigned*)out)[i] = (unsigned)f64_to_i64_x86(x); break;
    case AGC_DT_I16: case AGC_DT_U16: {
      int v = (!(x > -2147483649.0 && x < 2147483648.0)) ? (int)0x80000000 : (int)x;
      ((unsigned short*)out)[i] = (unsigned short)v; break; }
    default: {   // I8 / U8
      int v = (!(x > -2147483649.0 && x < 2147483648.0)) ? (int)0x80000000 : (int)x;
      ((unsigned char*)out)[i] = (unsigned char)v; break; }
  }
}
__device__ __forceinline__ void store_from_i64(void* out, int dt, long long i, long long x, bool is_unsigned = false) {
  switch (dt) {
    case AGC_DT_F64: ((double*)out)[i] = is_unsigned ? (double)(unsigned long long)x : (double)x; break;
    case AGC_DT_F32: ((float*)out)[i] = is_unsigned ? (float)(unsigned long long)x : (float)x; break;
    case AGC_DT_BOOL: ((unsigned char*)out)[i] = (x != 0); break;
    case AGC_DT_I64: case AGC_DT_U64: ((long long*)out)[i] = x; break;
    case AGC_DT_I32: case AGC_DT_U32: ((int*)out)[i] = (int)x; break;
    case AGC_DT_I16: case AGC_DT_U16: ((short*)out)[i] = (short)x; break;
    default: ((signed char*)out)[i] = (signed char)x; break;
  }
}

// streaming 128-bit load that does not allocate in L1 (inputs are read exactly once)
__device__ __forceinline__ int4 ldg_stream16(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// L2 eviction policy for data that is read exactly once: lets the streamed inputs leave L2 first so the
// group tables (up to ~100 MB) stay resident while 12.9 GB of labels/values pass through.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ int4 ldg_stream16_ef(const void* p, unsigned long long pol) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}

// RED with an L2 evict-last hint: keeps a large accumulator table (tens of MB) resident while GBs of input stream by
__device__ __forceinline__ unsigned long long l2_evict_last_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void red_add_f64_hint(double* addr, double v, unsigned long long pol) {
  asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" :: "l"(addr), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_u64_hint(unsigned long long* addr, unsigned long long v, unsigned long long pol) {
  asm volatile("red.global.add.L2::cache_hint.u64 [%0], %1, %2;" :: "l"(addr), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_max_s64_hint(long long* addr, long long v, unsigned long long pol) {
  asm volatile("red.global.max.L2::cache_hint.s64 [%0], %1, %2;" :: "l"(addr), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_min_s64_hint(long long* addr, long long v, unsigned long long pol) {
  asm volatile("red.global.min.L2::cache_hint.s64 [%0], %1, %2;" :: "l"(addr), "l"(v), "l"(pol) : "memory");
}

// shared-memory RED (no return value) on a 32-bit shared-window address
__device__ __forceinline__ void reds_add_u32(unsigned saddr, unsigned v) {
  asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}

}  // namespace agc
