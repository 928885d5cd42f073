// agc_generic.cuh -- the general scatter-reduce path: every op, every dtype, every Form,
// accumulator tables in global memory (L2-resident for realistic sizes), native RED atomics.
// The specialised shared-memory kernels in agc_fast.cuh take over when they apply; this file
// is the semantic ground truth on the device.
//
// Accumulator tables (SoA, `size` entries each, carved from the caller's workspace):
//   T0  8 B   f64 sum / i64 sum / f64|i64 product / min-max key / (var: mean after pass 1)
//   T1  8 B   i64 count  /  first|last|arg element index
//   T2  8 B   f64 sum of squared deviations (var/std second pass)
//   B0  1 B   any/all value      B1  1 B   "group was touched"
#pragma once
#include "agc_common.cuh"

namespace agc {

struct Tables {
  double* t0f; long long* t0i; long long* t1; double* t2; unsigned char* b0; unsigned char* b1;
};

struct AccParams {
  int op, flags, a_dtype, out_dtype, pass;
  const void* a; double a_sf; long long a_si;
  Indexer ix;
  long long n, size, index_base;
  Tables t;
  int* status;
};

struct FinParams {
  int op, flags, a_dtype, out_dtype;
  const void* a;
  void* out;
  double fill_f, ddof; long long fill_i;
  long long size, index_base, arg_stride, arg_len, n;
  Tables t;
};

template <class T> __device__ __forceinline__ T load_val(const AccParams& p, long long i);
template <> __device__ __forceinline__ float load_val<float>(const AccParams& p, long long i) {
  return p.a ? ((const float*)p.a)[i] : (float)p.a_sf;
}
template <> __device__ __forceinline__ double load_val<double>(const AccParams& p, long long i) {
  return p.a ? ((const double*)p.a)[i] : p.a_sf;
}
template <> __device__ __forceinline__ long long load_val<long long>(const AccParams& p, long long i) {
  return p.a ? load_int(p.a, p.a_dtype, i) : p.a_si;
}
template <> __device__ __forceinline__ unsigned long long load_val<unsigned long long>(const AccParams& p, long long i) {
  return p.a ? ((const unsigned long long*)p.a)[i] : (unsigned long long)p.a_si;
}

__device__ __forceinline__ void atomic_mul_f64(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)assumed) * v));
  } while (old != assumed);
}
__device__ __forceinline__ void atomic_mul_u64(unsigned long long* a, unsigned long long v) {
  unsigned long long old = *a, assumed;
  do { assumed = old; old = atomicCAS(a, assumed, assumed * v); } while (old != assumed);
}

template <class T> __device__ __forceinline__ double to_f64(T v) { return (double)v; }

// whether T0 holds a double (true) or an int64 (false) for the additive / multiplicative ops
__host__ __device__ inline bool acc_is_float(int op, int a_dtype, int out_dtype) {
  if (dtype_is_float(a_dtype)) return true;
  if (op == AGC_OP_PROD || op == AGC_OP_CUMPROD) return dtype_is_float(out_dtype);
  return false;
}

// ---------------------------------------------------------------------------------------------
// table initialisation
// ---------------------------------------------------------------------------------------------
__global__ void k_init_tables(Tables t, long long size, int op, int a_dtype, int out_dtype, int pass2_only) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const bool fl = acc_is_float(op, a_dtype, out_dtype);
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < size; g += stride) {
    if (pass2_only) {                          // second pass of var / arg*: only T2 / T1 are new
      if (op == AGC_OP_VAR) t.t2[g] = 0.0; else t.t1[g] = AGC_NO_INDEX;
      continue;
    }
    switch (op) {
      case AGC_OP_SUM: case AGC_OP_SUMSQ: t.t0i[g] = 0; t.b1[g] = 0; break;        // +0.0 == int 0 bit pattern
      case AGC_OP_PROD: if (fl) t.t0f[g] = 1.0; else t.t0i[g] = 1; t.b1[g] = 0; break;
      case AGC_OP_LEN: t.t1[g] = 0; break;
      case AGC_OP_MEAN: t.t0i[g] = 0; t.t1[g] = 0; break;
      case AGC_OP_VAR: t.t0i[g] = 0; t.t1[g] = 0; t.t2[g] = 0.0; break;
      case AGC_OP_MAX: case AGC_OP_ARGMAX: t.t0i[g] = AGC_EMPTY_MAX; t.t1[g] = AGC_NO_INDEX; t.b1[g] = 0; break;
      case AGC_OP_MIN: case AGC_OP_ARGMIN: t.t0i[g] = AGC_EMPTY_MIN; t.t1[g] = AGC_NO_INDEX; t.b1[g] = 0; break;
      case AGC_OP_MINMAX: t.t0i[g] = AGC_EMPTY_MAX; ((long long*)t.t2)[g] = AGC_EMPTY_MIN; t.b1[g] = 0; break;   // max keys in T0, min keys in T2
      case AGC_OP_ANY: case AGC_OP_ANYNAN: t.b0[g] = 0; t.b1[g] = 0; break;
      case AGC_OP_ALL: case AGC_OP_ALLNAN: t.b0[g] = 1; t.b1[g] = 0; break;
      case AGC_OP_FIRST: t.t1[g] = AGC_NO_INDEX; break;
      case AGC_OP_LAST: t.t1[g] = -1; break;
      default: break;
    }
  }
}

// var/std between the passes: T0 <- sum / count  (aggregate_numpy.py:180-183 `means = sums / counts`)
__global__ void k_means_inplace(Tables t, long long size) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < size; g += stride)
    t.t0f[g] = t.t0f[g] / (double)t.t1[g];
}

// ---------------------------------------------------------------------------------------------
// accumulate: one element -> its group's accumulators
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void accumulate_one(const AccParams& p, long long g, long long i, T v, unsigned long long pol) {
  const Tables& t = p.t;
  const bool nan = is_nan(v);
  if ((p.flags & AGC_F_NANSKIP) && nan) return;
  constexpr bool FL = (sizeof(T) == 4 && T(0.5f) != T(0)) || (sizeof(T) == 8 && T(0.5) != T(0));   // floating class
  if (p.pass == 2) {
    if (p.op == AGC_OP_VAR) {                       // sum (a - mean)^2 in f64, as the reference does
      double d = to_f64(v) - t.t0f[g];
      red_add_f64_hint(t.t2 + g, d * d, pol);
    } else {                                        // arg*: first element equal to the group's extreme
      if (!nan && key_of(v) == t.t0i[g] && i + p.index_base < t.t1[g]) atomicMin(t.t1 + g, i + p.index_base);   // stale read: only an extra atomic
    }
    return;
  }
  switch (p.op) {
    case AGC_OP_SUM:
      if (FL) red_add_f64_hint(t.t0f + g, to_f64(v), pol);
      else red_add_u64_hint((unsigned long long*)(t.t0i + g), (unsigned long long)v, pol);
      if ((p.flags & AGC_F_NEED_TOUCHED) && !t.b1[g]) t.b1[g] = 1;
      break;
    case AGC_OP_SUMSQ:
      if (FL) atomicAdd(t.t0f + g, to_f64((T)(v * v)));
      else atomicAdd((unsigned long long*)(t.t0i + g),
                     (unsigned long long)wrap_int((long long)((unsigned long long)v * (unsigned long long)v), p.a_dtype));
      if ((p.flags & AGC_F_NEED_TOUCHED) && !t.b1[g]) t.b1[g] = 1;
      break;
    case AGC_OP_PROD:
      if (acc_is_float(p.op, p.a_dtype, p.out_dtype)) atomic_mul_f64(t.t0f + g, to_f64(v));
      else atomic_mul_u64((unsigned long long*)(t.t0i + g), (unsigned long long)v);
      if ((p.flags & AGC_F_NEED_TOUCHED) && !t.b1[g]) t.b1[g] = 1;
      break;
    case AGC_OP_LEN:
      red_add_u64_hint((unsigned long long*)(t.t1 + g), 1ull, pol);
      break;
    case AGC_OP_MEAN: case AGC_OP_VAR:
      red_add_f64_hint(t.t0f + g, to_f64(v), pol);
      red_add_u64_hint((unsigned long long*)(t.t1 + g), 1ull, pol);
      break;
    case AGC_OP_MAX: case AGC_OP_ARGMAX: {
      long long k = nan ? AGC_EMPTY_MIN /* = INT64_MAX: NaN beats everything */ : key_of(v);
      // monotone update: a plain (possibly stale) read can only make us issue an unnecessary RED, never skip a needed
      // one -- and with few groups almost every element fails the test, so the hot L2 addresses see no traffic
      if (k > t.t0i[g]) red_max_s64_hint(t.t0i + g, k, pol);
      if (!FL && !t.b1[g]) t.b1[g] = 1;
      break; }
    case AGC_OP_MIN: case AGC_OP_ARGMIN: {
      long long k = nan ? AGC_EMPTY_MAX /* = INT64_MIN */ : key_of(v);
      if (k < t.t0i[g]) red_min_s64_hint(t.t0i + g, k, pol);
      if (!FL && !t.b1[g]) t.b1[g] = 1;
      break; }
    case AGC_OP_MINMAX: {                            // fused multi-statistic pass: both extremes from one read of (labels, a)
      const long long kmax = nan ? AGC_EMPTY_MIN : key_of(v), kmin = nan ? AGC_EMPTY_MAX : key_of(v);
      long long* tmin = (long long*)t.t2;
      if (kmax > t.t0i[g]) red_max_s64_hint(t.t0i + g, kmax, pol);
      if (kmin < tmin[g]) red_min_s64_hint(tmin + g, kmin, pol);
      if (!FL && !t.b1[g]) t.b1[g] = 1;
      break; }
    case AGC_OP_ANY: if (v != T(0) && !t.b0[g]) t.b0[g] = 1; if (!t.b1[g]) t.b1[g] = 1; break;       // NaN is truthy; test-then-set keeps hot bytes read-only
    case AGC_OP_ALL: if (v == T(0) && t.b0[g]) t.b0[g] = 0; if (!t.b1[g]) t.b1[g] = 1; break;
    case AGC_OP_ANYNAN: if (nan && !t.b0[g]) t.b0[g] = 1; if (!t.b1[g]) t.b1[g] = 1; break;
    case AGC_OP_ALLNAN: if (!nan && t.b0[g]) t.b0[g] = 0; if (!t.b1[g]) t.b1[g] = 1; break;
    case AGC_OP_FIRST: if (i + p.index_base < t.t1[g]) atomicMin(t.t1 + g, i + p.index_base); break;
    case AGC_OP_LAST: if (i + p.index_base > t.t1[g]) atomicMax(t.t1 + g, i + p.index_base); break;
    default: break;
  }
}

template <class T>
__global__ void __launch_bounds__(256) k_accumulate(const AccParams p) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const unsigned long long pol = l2_evict_last_policy();      // accumulator tables stay in L2 while the inputs stream by
  bool bad = false;
  const bool backwards = p.op == AGC_OP_LAST;                 // `last` walks from the end: after the first wave no element improves
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < p.n; j += stride) {
    const long long i = backwards ? p.n - 1 - j : j;
    long long g = p.ix.group(i);
    if (g < 0) { bad = true; continue; }
    T v = (p.op == AGC_OP_LEN && !(p.flags & AGC_F_NANSKIP)) ? T(0) : load_val<T>(p, i);
    accumulate_one<T>(p, g, i, v, pol);
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
}

// ---------------------------------------------------------------------------------------------
// finalise: accumulators -> output dtype, fill_value for groups that never saw an element
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fin_store_fill(const FinParams& p, long long g) {
  if (dtype_is_float(p.out_dtype)) store_from_f64(p.out, p.out_dtype, g, p.fill_f);
  else store_from_i64(p.out, p.out_dtype, g, p.fill_i);
}

__global__ void k_finalize(const FinParams p) {
  const Tables& t = p.t;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool a_fl = dtype_is_float(p.a_dtype);
  const bool a_u64 = p.a_dtype == AGC_DT_U64;
  const bool fl = acc_is_float(p.op, p.a_dtype, p.out_dtype);
  const bool need_touched = p.flags & AGC_F_NEED_TOUCHED;
  const bool fill_nan = p.flags & AGC_F_FILL_NAN;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < p.size; g += stride) {
    switch (p.op) {
      case AGC_OP_SUM: case AGC_OP_SUMSQ: case AGC_OP_PROD:
        if (need_touched && ((p.flags & AGC_F_TOUCHED_BY_COUNT) ? t.t1[g] == 0 : !t.b1[g])) fin_store_fill(p, g);
        else if (fl) store_from_f64(p.out, p.out_dtype, g, t.t0f[g]);
        else store_from_i64(p.out, p.out_dtype, g, t.t0i[g], a_u64);
        break;
      case AGC_OP_LEN: {
        long long c = t.t1[g];
        if (c == 0 && need_touched) fin_store_fill(p, g); else store_from_i64(p.out, p.out_dtype, g, c);
        break; }
      case AGC_OP_MEAN: {
        long long c = t.t1[g];
        if (c == 0 && !fill_nan) fin_store_fill(p, g);
        else store_from_f64(p.out, p.out_dtype, g, t.t0f[g] / (double)c);
        break; }
      case AGC_OP_VAR: {
        double c = (double)t.t1[g];
        double d = c > p.ddof ? c - p.ddof : 0.0;
        double r = t.t2[g] / d;
        if (p.flags & AGC_F_SQRT) r = sqrt(r);
        if (!fill_nan && d == 0.0) fin_store_fill(p, g); else store_from_f64(p.out, p.out_dtype, g, r);
        break; }
      case AGC_OP_MAX: case AGC_OP_MIN: {
        long long k = (p.flags & AGC_F_KEYS_IN_T2) ? ((const long long*)t.t2)[g] : t.t0i[g];
        const long long empty = p.op == AGC_OP_MAX ? AGC_EMPTY_MAX : AGC_EMPTY_MIN;
        const long long nankey = p.op == AGC_OP_MAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX;
        if (a_fl) {
          if (k == empty) fin_store_fill(p, g);
          else store_from_f64(p.out, p.out_dtype, g, k == nankey ? __longlong_as_double(0x7ff8000000000000ll) : unkey_f64(k));
        } else {
          if (!t.b1[g]) fin_store_fill(p, g);
          else store_from_i64(p.out, p.out_dtype, g, a_u64 ? (long long)((unsigned long long)k ^ 0x8000000000000000ull) : k, a_u64);
        }
        break; }
      case AGC_OP_ANY: case AGC_OP_ALL: case AGC_OP_ANYNAN: case AGC_OP_ALLNAN:
        store_from_i64(p.out, p.out_dtype, g, t.b1[g] ? (long long)t.b0[g] : (long long)(p.fill_i != 0));
        break;
      case AGC_OP_FIRST: case AGC_OP_LAST: {
        long long idx = t.t1[g];
        if (idx == AGC_NO_INDEX || idx < 0) { fin_store_fill(p, g); break; }
        idx -= p.index_base;
        if (idx < 0 || idx >= p.n) { fin_store_fill(p, g); break; }     // owned by another shard
        if (p.a_dtype == AGC_DT_F32) store_from_f64(p.out, p.out_dtype, g, (double)((const float*)p.a)[idx]);
        else if (p.a_dtype == AGC_DT_F64) store_from_f64(p.out, p.out_dtype, g, ((const double*)p.a)[idx]);
        else store_from_i64(p.out, p.out_dtype, g, load_int(p.a, p.a_dtype, idx), a_u64);
        break; }
      case AGC_OP_ARGMAX: case AGC_OP_ARGMIN: {
        long long idx = t.t1[g];
        if (idx == AGC_NO_INDEX) fin_store_fill(p, g);
        else store_from_i64(p.out, p.out_dtype, g, p.arg_len > 0 ? (idx / p.arg_stride) % p.arg_len : idx);
        break; }
      default: break;
    }
  }
}

}  // namespace agc
