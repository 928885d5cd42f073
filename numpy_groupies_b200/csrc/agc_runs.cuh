// agc_runs.cuh -- reductions over labels that come in RUNS (sorted or contiguous group_idx) when the table is too
// large for shared memory: a segmented warp reduce instead of one atomic per element.
//
// A thread takes 4 consecutive elements (one 128-bit value load, 4 labels).  If its 4 labels are equal -- the normal
// case inside a run -- it folds them into one (label, partial) pair; the 32 pairs of a warp then go through a
// segmented inclusive scan by label (5 shuffle steps) and the LAST lane of every segment sends one RED to the global
// accumulator tables of agc_generic.cuh.  A thread that straddles a run boundary sends its elements one by one and
// acts as a segment break.  For runs of >= a few hundred elements that is 1-2 REDs per 128 elements, so the kernel
// streams at HBM rate; on random labels it degenerates into the plain global-RED path (never selected there: the
// device-side label probe routes only `equal neighbours > 50 %` here).  Partials are f64 (sums of f32 values),
// int counts and order-preserving s32 keys; the result does not depend on how runs fall on warps beyond f64 rounding.
#pragma once
#include "agc_fast.cuh"

namespace agc {

template <int MODE, class LT, bool LOAD_A>
__global__ void __launch_bounds__(FAST_THREADS, 2) k_runs(const FastParams p) {
  if (p.choose && *p.choose != p.want) return;
  constexpr bool IS_SUM = MODE == FM_SUM || MODE == FM_SUMCNT;
  constexpr bool HAS_CNT = MODE == FM_SUMCNT || MODE == FM_LEN;
  const int lane = threadIdx.x & 31;
  const long long nvec = p.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  const int ident = MODE == FM_MAX ? (int)0x80000000 : 0x7fffffff;          // "no element" key (a NaN pattern, never a real key)
  const unsigned long long pol = l2_evict_last_policy();
  bool bad = false;

  // one element straight to the global tables
  auto emit1 = [&](long long g, float x) {
    const bool nan = x != x;
    if (LOAD_A && nan && nanskip) return;
    if (MODE == FM_LEN) atomicAdd((unsigned long long*)(p.t.t1 + g), 1ull);
    else if (MODE == FM_MAX || MODE == FM_MIN) {
      const long long k64 = nan ? (MODE == FM_MAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX) : key_of((double)x);
      if (MODE == FM_MAX) atomicMax(p.t.t0i + g, k64); else atomicMin(p.t.t0i + g, k64);
    } else {
      if (p.square) x = x * x;
      if ((__float_as_uint(x) & 0x7fffffffu) != 0u) atomicAdd(p.t.t0f + g, (double)x);
      if (MODE == FM_SUMCNT) { if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + g), 1ull); if (p.want_b1) p.t.b1[g] = 1; }
    }
  };

  // warp-uniform trip count: every lane takes part in the shuffles of every iteration
  for (long long vb = ((long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31)); vb < nvec; vb += stride) {
    const long long v = vb + lane;
    const bool in = v < nvec;
    long long g[4] = {-1, -1, -1, -1};
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (in) {
      LabelVec<LT>::load(p.labels, v, g);
      if (LOAD_A) {
        int4 av = ldg_stream16((const int4*)p.a + v);
        x[0] = __int_as_float(av.x); x[1] = __int_as_float(av.y); x[2] = __int_as_float(av.z); x[3] = __int_as_float(av.w);
      }
    }
    const bool ok0 = in && (unsigned long long)g[0] < (unsigned long long)p.size;
    const bool uniform = ok0 && g[1] == g[0] && g[2] == g[0] && g[3] == g[0];
    double s = 0.0; int c = 0; int k = ident;
    if (uniform) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float xv = x[e];
        const bool nan = xv != xv;
        if (LOAD_A && nan && nanskip) continue;
        if (IS_SUM) { if (p.square) xv = xv * xv; s += (double)xv; }
        if (HAS_CNT) ++c;
        if (MODE == FM_MAX || MODE == FM_MIN) {
          const int ke = nan ? (MODE == FM_MAX ? 0x7fffffff : (int)0x80000000) : key32_of(xv);     // NaN wins
          k = MODE == FM_MAX ? max(k, ke) : min(k, ke);
        }
      }
    } else if (in) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if ((unsigned long long)g[e] >= (unsigned long long)p.size) bad = true; else emit1(g[e], x[e]);
      }
    }
    // segmented inclusive scan over the lanes that hold one (label, partial) pair
    const long long lab = uniform ? g[0] : -1 - lane;                    // non-uniform lanes never match a neighbour
    const long long plab = __shfl_up_sync(0xffffffffu, lab, 1);
    bool f = lane == 0 || plab != lab;                                    // segment head
    const bool head = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tf = __shfl_up_sync(0xffffffffu, (int)f, o);
      double ts = 0.0; int tc = 0, tk = ident;
      if (IS_SUM) ts = __shfl_up_sync(0xffffffffu, s, o);
      if (HAS_CNT) tc = __shfl_up_sync(0xffffffffu, c, o);
      if (MODE == FM_MAX || MODE == FM_MIN) tk = __shfl_up_sync(0xffffffffu, k, o);
      if (lane >= o && !f) {
        if (IS_SUM) s += ts;
        if (HAS_CNT) c += tc;
        if (MODE == FM_MAX) k = max(k, tk);
        if (MODE == FM_MIN) k = min(k, tk);
        f = tf != 0;
      }
    }
    const int nhead = __shfl_down_sync(0xffffffffu, (int)head, 1);
    const bool last = uniform && (lane == 31 || nhead);
    if (last) {
      const long long gg = g[0];
      if (MODE == FM_LEN) { if (c) red_add_u64_hint((unsigned long long*)(p.t.t1 + gg), (unsigned long long)c, pol); }
      else if (MODE == FM_MAX || MODE == FM_MIN) {
        if (k != ident) {
          long long k64;
          if (MODE == FM_MAX) k64 = k == 0x7fffffff ? AGC_EMPTY_MIN : key_of((double)unkey_f32(k));
          else k64 = k == (int)0x80000000 ? AGC_EMPTY_MAX : key_of((double)unkey_f32(k));
          if (MODE == FM_MAX) red_max_s64_hint(p.t.t0i + gg, k64, pol); else red_min_s64_hint(p.t.t0i + gg, k64, pol);
        }
      } else {
        if (s != 0.0) red_add_f64_hint(p.t.t0f + gg, s, pol);
        if (MODE == FM_SUMCNT && c) {
          if (p.want_cnt) red_add_u64_hint((unsigned long long*)(p.t.t1 + gg), (unsigned long long)c, pol);
          if (p.want_b1) p.t.b1[gg] = 1;
        }
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {                      // tail elements
    const long long i = (nvec << 2) + threadIdx.x;
    const long long gg = LabelVec<LT>::one(p.labels, i);
    if ((unsigned long long)gg >= (unsigned long long)p.size) bad = true; else emit1(gg, LOAD_A ? p.a[i] : 0.f);
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.status[AGC_ST_REGIME] = 50 + MODE;
}

template <int MODE, class LT, bool LOAD_A>
inline int launch_runs(const FastParams& p, cudaStream_t st, int sms) {
  AGC_COUNT_LAUNCH(1), k_runs<MODE, LT, LOAD_A><<<sms * 2, FAST_THREADS, 0, st>>>(p);
  return 1;
}

}  // namespace agc
