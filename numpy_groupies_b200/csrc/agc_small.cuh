// agc_small.cuh -- the general path for SMALL tables: additive / multiplicative ops on any value dtype and any Form.
//
// With a few hundred groups every global RED of k_accumulate lands on the same handful of L2 addresses and the
// whole GPU serialises behind them (measured at 100 groups, N = 2^28: f64 sum 6.0, int64 sum 6.8, mean 3.1,
// prod 2.0 Gelem/s).  Here every CTA keeps a private copy of the accumulators in shared memory, replicated per lane
// (slot = group * R + lane % R) so that a warp never collides with itself; 64-bit shared atomics are CAS loops on
// sm_100a, but with replication a CAS almost always succeeds at the first try.  One RED per group and CTA merges the
// copies into the global tables of agc_generic.cuh, so finalise / multi-GPU / second passes are unchanged.
// Ops: sum, sumofsquares, mean, var (both passes), prod.  (min / max / arg* / first / last / any / all stay on
// k_accumulate: their updates are monotone and test-then-update already keeps the hot addresses read-only.)
#pragma once
#include "agc_generic.cuh"

namespace agc {

constexpr int SMALL_THREADS = 512;
constexpr size_t SMALL_SMEM = 47 * 1024;        // 4 CTAs/SM, replicated tables (4 x (47 + 1) KB <= 196 KB per SM)
constexpr size_t SMALL_SMEM_BIG = 194 * 1024;   // up to ~16.5 K groups: one unreplicated table per SM (<= 196 KB per SM keeps 60 KB of L1)

__device__ __forceinline__ void smem_mul_f64(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)assumed) * v));
  } while (old != assumed);
}
__device__ __forceinline__ void smem_mul_u64(unsigned long long* a, unsigned long long v) {
  unsigned long long old = *a, assumed;
  do { assumed = old; old = atomicCAS(a, assumed, assumed * v); } while (old != assumed);
}

inline bool small_path_ok(int op, int pass, long long n, long long size) {
  if (!(op == AGC_OP_SUM || op == AGC_OP_SUMSQ || op == AGC_OP_MEAN || op == AGC_OP_VAR || op == AGC_OP_PROD)) return false;
  (void)pass;
  return n >= (1 << 16) && size > 0 && (size_t)size * 12 <= SMALL_SMEM_BIG;
}

template <class T>
__global__ void __launch_bounds__(1024) k_accumulate_small(const AccParams p, int rshift) {
  extern __shared__ __align__(16) unsigned char small_smem[];
  constexpr bool FL = (sizeof(T) == 4 && T(0.5f) != T(0)) || (sizeof(T) == 8 && T(0.5) != T(0));   // floating class
  const int S = (int)p.size << rshift;
  double* s0f = (double*)small_smem; unsigned long long* s0i = (unsigned long long*)small_smem;
  unsigned* s1 = (unsigned*)(small_smem + (size_t)S * 8);              // element counts (mean / var / touched)
  const bool facc = acc_is_float(p.op, p.a_dtype, p.out_dtype);
  const bool is_prod = p.op == AGC_OP_PROD;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    if (is_prod) { if (facc) s0f[i] = 1.0; else s0i[i] = 1ull; } else s0i[i] = 0ull;     // +0.0 == integer 0
    s1[i] = 0u;
  }
  __syncthreads();
  const int rsel = (threadIdx.x & 31) & ((1 << rshift) - 1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  bool bad = false;
  auto acc = [&](long long g, T v) {
    if (nanskip && is_nan(v)) return;
    const int slot = ((int)g << rshift) | rsel;
    if (p.pass == 2) {                                    // var / std: sum (a - mean)^2, means already in T0
      const double d = to_f64(v) - p.t.t0f[g];
      atomicAdd(s0f + slot, d * d);
      atomicAdd(s1 + slot, 1u);
      return;
    }
    switch (p.op) {
      case AGC_OP_SUM:
        if (FL) atomicAdd(s0f + slot, to_f64(v)); else atomicAdd(s0i + slot, (unsigned long long)v);
        break;
      case AGC_OP_SUMSQ:
        if (FL) atomicAdd(s0f + slot, to_f64((T)(v * v)));
        else atomicAdd(s0i + slot, (unsigned long long)wrap_int((long long)((unsigned long long)v * (unsigned long long)v), p.a_dtype));
        break;
      case AGC_OP_PROD:
        if (facc) smem_mul_f64(s0f + slot, to_f64(v)); else smem_mul_u64(s0i + slot, (unsigned long long)v);
        break;
      default:                                            // mean / var pass 1
        atomicAdd(s0f + slot, to_f64(v));
        break;
    }
    atomicAdd(s1 + slot, 1u);
  };
  // Form 1, int64 labels, 8-byte values (f64 / int64 / uint64), 16-byte aligned: 128-bit streaming loads, two vectors
  // (4 elements) per thread in flight -- the generic loop below costs a runtime dtype / Form switch per element
  const bool fast8 = p.ix.ix.form == AGC_FORM_FLAT && p.ix.ix.dtype == AGC_DT_I64 && p.a != nullptr && sizeof(T) == 8 &&
                     dtype_bytes(p.a_dtype) == 8 && !((uintptr_t)p.ix.ix.labels & 15) && !((uintptr_t)p.a & 15);
  long long i0 = 0;
  if (fast8) {
    const long long nv = p.n >> 1;                        // vectors of 2 elements
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nv; v += 2 * stride) {
      const long long v2 = v + stride;
      const bool two = v2 < nv;
      const int4 l0 = ldg_stream16((const int4*)p.ix.ix.labels + v), a0 = ldg_stream16((const int4*)p.a + v);
      int4 l1 = make_int4(0, 0, 0, 0), a1 = l1;
      if (two) { l1 = ldg_stream16((const int4*)p.ix.ix.labels + v2); a1 = ldg_stream16((const int4*)p.a + v2); }
      const int4 ls[2] = {l0, l1}, as[2] = {a0, a1};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const long long g0 = ((long long)(unsigned)ls[u].x) | ((long long)ls[u].y << 32), g1 = ((long long)(unsigned)ls[u].z) | ((long long)ls[u].w << 32);
        const long long b0 = ((long long)(unsigned)as[u].x) | ((long long)as[u].y << 32), b1 = ((long long)(unsigned)as[u].z) | ((long long)as[u].w << 32);
        T x0, x1;
        if (FL) { x0 = (T)__longlong_as_double(b0); x1 = (T)__longlong_as_double(b1); } else { x0 = (T)b0; x1 = (T)b1; }
        if ((unsigned long long)g0 >= (unsigned long long)p.size) bad = true; else acc(g0, x0);
        if ((unsigned long long)g1 >= (unsigned long long)p.size) bad = true; else acc(g1, x1);
      }
    }
    i0 = nv << 1;                                         // odd tail element: generic loop
  }
  for (long long i = i0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n; i += stride) {
    const long long g = p.ix.group(i);
    if (g < 0) { bad = true; continue; }
    acc(g, load_val<T>(p, i));
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.pass == 1) p.status[AGC_ST_REGIME] = 60;
  __syncthreads();
  const int R = 1 << rshift;
  for (int g = threadIdx.x; g < (int)p.size; g += blockDim.x) {
    unsigned long long c = 0;
    for (int r = 0; r < R; ++r) c += s1[(g << rshift) + r];
    if (c == 0) continue;
    if (p.pass == 2) {
      double s = 0.0;
      for (int r = 0; r < R; ++r) s += s0f[(g << rshift) + r];
      atomicAdd(p.t.t2 + g, s);
      continue;
    }
    if (is_prod) {
      if (facc) { double m = 1.0; for (int r = 0; r < R; ++r) m *= s0f[(g << rshift) + r]; atomic_mul_f64(p.t.t0f + g, m); }
      else { unsigned long long m = 1ull; for (int r = 0; r < R; ++r) m *= s0i[(g << rshift) + r]; atomic_mul_u64((unsigned long long*)(p.t.t0i + g), m); }
      if (p.flags & AGC_F_NEED_TOUCHED) p.t.b1[g] = 1;
    } else if (p.op == AGC_OP_SUM || p.op == AGC_OP_SUMSQ) {
      if (FL) { double s = 0.0; for (int r = 0; r < R; ++r) s += s0f[(g << rshift) + r]; atomicAdd(p.t.t0f + g, s); }
      else { unsigned long long s = 0ull; for (int r = 0; r < R; ++r) s += s0i[(g << rshift) + r]; atomicAdd((unsigned long long*)(p.t.t0i + g), s); }
      if (p.flags & AGC_F_NEED_TOUCHED) p.t.b1[g] = 1;
    } else {
      double s = 0.0; for (int r = 0; r < R; ++r) s += s0f[(g << rshift) + r];
      atomicAdd(p.t.t0f + g, s);
      atomicAdd((unsigned long long*)(p.t.t1 + g), c);
    }
  }
}

template <class T>
inline void launch_accumulate_small(const AccParams& p, cudaStream_t st, int sms) {
  if ((size_t)p.size * 12 > SMALL_SMEM) {                 // one big table per SM, two CTAs of 512 threads cannot both hold it
    static std::atomic<unsigned long long> attr_done{0};
    if (!ensure_dyn_smem(k_accumulate_small<T>, (int)SMALL_SMEM_BIG, attr_done)) {       // cannot take the big table: general kernel
      AGC_COUNT_LAUNCH(1), k_accumulate<T><<<tile_grid((p.n + 255) / 256, 16), 256, 0, st>>>(p);
      return;
    }
    long long blocks = (p.n + 1024 * 8 - 1) / (1024 * 8);
    if (blocks > sms) blocks = sms;
    AGC_COUNT_LAUNCH(1), k_accumulate_small<T><<<(int)blocks, 1024, (size_t)p.size * 12, st>>>(p, 0);
    return;
  }
  int rshift = 0;
  while (rshift < 5 && ((size_t)p.size << (rshift + 1)) * 12 <= SMALL_SMEM) ++rshift;
  const size_t smem = ((size_t)p.size << rshift) * 12;
  long long blocks = (p.n + SMALL_THREADS * 8 - 1) / (SMALL_THREADS * 8);
  const long long cap = (long long)sms * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  AGC_COUNT_LAUNCH(1), k_accumulate_small<T><<<(int)blocks, SMALL_THREADS, smem, st>>>(p, rshift);
}

}  // namespace agc
