// agc_fast.cuh -- shared-memory-privatised scatter-reduce for the headline Form-1 workloads
// (f32 values, int64/int32 labels; sum / sumofsquares / mean / min / max / len and nan-variants).
//
// Design, from the measurements in ubench/ (profiles/r01_probe*.md):
//   * the only shared-memory atomics that are native on sm_100a are 32-bit INTEGER ops
//     (ATOMS.ADD / ATOMS.MIN / ATOMS.MAX); float atomicAdd in shared memory is a CAS loop that
//     collapses under skew, global RED collapses on hot addresses.  Native ATOMS at 2048
//     threads/SM sustain ~500 Gelem/s (>90 % of the HBM roofline) for uniform AND Zipf labels.
//   * so every accumulator is an integer:
//       sum   -> 64-bit FIXED POINT (two u32 words, carry propagated with a second ATOMS) whose
//                scale is chosen per CTA from the data: a value whose exponent lies in the
//                16-binade window below the CTA's reference exponent converts EXACTLY
//                (24-bit mantissa << 1..16); anything else (huge, tiny, inf, NaN) bypasses to the
//                global f64 table with one RED.  The result is therefore at least as accurate
//                as the reference's sequential f64 sum, and independent of the order of
//                elements inside a CTA.
//       count -> u32 word;   min/max -> order-preserving s32 key of the f32 bits.
//   * each CTA owns a private table in shared memory (size*words*4 B), streams its share of
//     (labels, a) with 128-bit L1-bypassing loads, and finally merges the table into the global
//     accumulator tables of agc_generic.cuh with one RED per touched group.
#pragma once
#include "agc_generic.cuh"

namespace agc {

enum { FM_SUM = 0, FM_SUMCNT = 1, FM_MAX = 2, FM_MIN = 3, FM_LEN = 4, FM_SUM1 = 5, FM_SUMCNT1 = 6, FM_MINMAX = 7 };
// FM_MINMAX: both extremes from one read of (labels, a) -- the fused multi-statistic pass (AGC_OP_MINMAX): max key in word 0,
// min key in word 1, two native ATOMS per element; merged into T0 (max keys) and T2 (min keys, as int64)
// FM_SUMCNT1: FM_SUM1's one-word sum + a u32 count (8 B per group instead of FM_SUMCNT's 12): mean / var on tables of 16.6 K - 24.9 K groups
// FM_SUM1: ONE 32-bit word per group (k_fast only).  The word starts at 2^31 and takes q = x * 2^(FX1_F - E) with a
// native ATOMS.ADD that returns the old value; a carry / borrow out of the word (the add crossed 0 or 2^32) is
// forwarded at once as one exact f64 RED of +-2^32 units to the global table, so no bit is ever lost and there is
// no per-launch headroom limit.  Half the shared memory per group: tables of up to 56 K groups fit one SM, and the
// partial table of larger sizes covers twice as many groups -- at 10^5 groups that is the difference between 72 %
// and ~48 % of the elements leaving the SM as RED packets.  The price is a 6-binade exact window (below it: bypass
// RED), which is why the small-table kernels keep the two-word format.
constexpr int FX1_F = 29;
constexpr int FX1_WINDOW = 6;     // FX1_F - 23

// A narrower window (FX_F = 31: |q| < 2^31, the high word then moves only on a carry, ~1.25 ATOMS per element)
// was measured and rejected: the 0.4 % of rand() values that fall below an 8-binade window take the global RED
// bypass, and with few groups those REDs pile onto ~100 hot L2 addresses (sum G=100: 551 -> 356 Gelem/s).
constexpr int FX_F = 40;          // fixed-point fraction bits below the reference exponent
constexpr int FX_WINDOW = 16;     // binades that convert exactly: FX_F - 24
constexpr int FAST_THREADS = 1024;
constexpr long long FAST_MAX_PER_CTA = 1ll << 22;   // keeps 64-bit fixed point and u32 counts from overflowing

struct FastParams {
  const void* labels; const float* a;
  long long n, size;
  Tables t;
  int* status;
  int flags;            // AGC_F_NANSKIP, AGC_F_NEED_TOUCHED
  int square;           // sumofsquares: accumulate (float)(v*v)
  int want_b1;          // write touched flags (sum with fill != 0)
  int want_cnt;         // merge counts into T1 (mean / var)
  int rshift;           // k_fast: log2(replicas of the table per CTA) -- kills same-address ATOMS conflicts under skew
  int nslots;           // k_cache: slots of the per-CTA hot-key cache
  int cov;              // k_fast: groups [0, cov) live in shared memory, the rest go straight to the global tables
  const int* choose;    // device word written by k_skew_probe; run iff *choose == want (NULL: always)
  int want;
  const double* means;  // second pass of var / std: accumulate (float)((a - means[g])^2) instead of a (NULL: off)
  int means_smem;       // ... with a copy of the means in shared memory behind the table (size * 8 more bytes)
  int side;             // k_fast / k_cache sums: an f64 side accumulator per table entry / slot (8 more bytes) takes what falls outside the fixed-point window
  const int* choose2;   // k_cache: second device condition (the probe's "values span more than the window" word); run iff *choose2 == want2
  int want2;
};

// one out-of-window value into the CTA's f64 side accumulator (shared memory, CAS loop) -- deliberately not inlined
__device__ __noinline__ void smem_side_add(unsigned* side_words, int gi, double v) { atomicAdd((double*)side_words + gi, v); }

__device__ __forceinline__ void fx_add(unsigned* lo, unsigned* hi, long long q) {
  unsigned ql = (unsigned)q; int qh = (int)(q >> 32);
  unsigned old = atomicAdd(lo, ql);
  qh += (old + ql < old);                       // carry out of the low word
  if (qh != 0) atomicAdd(hi, (unsigned)qh);
}

template <class LT> struct LabelVec;
template <> struct LabelVec<long long> {
  // 4 consecutive int64 labels = two 16-byte loads
  static __device__ __forceinline__ void load(const void* base, long long v, long long (&g)[4]) {
    int4 l0 = ldg_stream16((const int4*)base + 2 * v), l1 = ldg_stream16((const int4*)base + 2 * v + 1);
    g[0] = ((long long)(unsigned)l0.x) | ((long long)l0.y << 32); g[1] = ((long long)(unsigned)l0.z) | ((long long)l0.w << 32);
    g[2] = ((long long)(unsigned)l1.x) | ((long long)l1.y << 32); g[3] = ((long long)(unsigned)l1.z) | ((long long)l1.w << 32);
  }
  static __device__ __forceinline__ long long one(const void* base, long long i) { return ((const long long*)base)[i]; }
};
template <> struct LabelVec<int> {
  static __device__ __forceinline__ void load(const void* base, long long v, long long (&g)[4]) {
    int4 l = ldg_stream16((const int4*)base + v);
    g[0] = l.x; g[1] = l.y; g[2] = l.z; g[3] = l.w;
  }
  static __device__ __forceinline__ long long one(const void* base, long long i) { return ((const int*)base)[i]; }
};

// U = 128-bit value vectors per thread per iteration: U = 1 runs 2 CTAs/SM (32 registers), U = 2 runs 1 CTA/SM with
// twice the loads in flight per thread (tables between 110 and 226 KB, and the partial / paired modes).
// PARTIAL: the table covers only groups [0, cov) (size too large for shared memory); the rest are global REDs.
// X: the instantiation that knows about the f64 side accumulator and the second pass of var / std (p.side, p.means); the
// plain one carries none of that code (measured: the runtime checks alone cost the 64-register variants 6 %).
template <int MODE, class LT, bool LOAD_A, int U, bool PARTIAL, bool X>
__global__ void __launch_bounds__(FAST_THREADS, U == 1 ? 2 : 1) k_fast(const FastParams p) {
  if (p.choose && *p.choose != p.want) return;
  const double* const xmeans = X ? p.means : nullptr;
  const int xside = X ? p.side : 0;
  extern __shared__ unsigned sw[];
  __shared__ unsigned s_maxbits;
  constexpr bool ONEWORD = MODE == FM_SUM1 || MODE == FM_SUMCNT1;
  constexpr bool HAS_CNT = MODE == FM_SUMCNT || MODE == FM_SUMCNT1;
  constexpr int WORDS = (MODE == FM_SUM || MODE == FM_SUMCNT1 || MODE == FM_MINMAX) ? 2 : (MODE == FM_SUMCNT ? 3 : 1);
  constexpr bool IS_SUM = MODE == FM_SUM || MODE == FM_SUMCNT || ONEWORD;
  constexpr int FXF = ONEWORD ? FX1_F : FX_F, FXW = ONEWORD ? FX1_WINDOW : FX_WINDOW;
  const int rs = p.rshift;
  const int G = p.cov << rs;                       // table entries incl. replicas: entry = (g << rs) | (lane & (R-1))
  const int rsel = (threadIdx.x & 31) & ((1 << rs) - 1);
  unsigned* w0 = sw;            // sum.lo | key | count(FM_LEN)
  unsigned* w1 = sw + G;        // sum.hi | count (FM_SUMCNT1)
  unsigned* w2 = sw + 2 * G;    // count (FM_SUMCNT)
  unsigned* wc = MODE == FM_SUMCNT1 ? w1 : w2;
  const unsigned init0 = (MODE == FM_MAX || MODE == FM_MINMAX || ONEWORD) ? 0x80000000u : (MODE == FM_MIN ? 0x7fffffffu : 0u);
  for (int i = threadIdx.x; i < WORDS * G; i += blockDim.x) sw[i] = (i < G) ? init0 : (MODE == FM_MINMAX ? 0x7fffffffu : 0u);   // G counts replicas
  // second pass of var / std with the means in shared memory: each mean as a (hi, lo) pair of floats (exact to 2^-48), so
  // that the deviation is two f32 subtractions instead of f64 arithmetic (relative error ~1e-7 per term, unbiased)
  // values outside the exact fixed-point window (huge / tiny / inf / NaN): on a small table their f64 REDs would queue
  // on a few hot L2 addresses (log-uniform data over 20 binades, 100 groups: 21 ms instead of 0.5 ms at N = 2^28), so
  // while it fits the CTA keeps an f64 side accumulator per table entry in shared memory (CAS-loop adds, used rarely
  // for narrow data, and still >100 Gelem/s when half the data takes it)
  double* side = nullptr;
  int tail_words = (WORDS * G + 1) & ~1;                   // first free (8-byte aligned) word behind the table
  if (IS_SUM && xside) {
    side = (double*)(sw + tail_words);
    for (int i = threadIdx.x; i < G; i += blockDim.x) side[i] = 0.0;
    tail_words += 2 * G;
  }
  const float2* mpair = nullptr;
  if (IS_SUM && xmeans && p.means_smem) {
    float2* sm = (float2*)(sw + tail_words);
    for (int i = threadIdx.x; i < (int)p.size; i += blockDim.x) {
      const double m = xmeans[i];
      float2 hl; hl.x = (float)m; hl.y = (float)(m - (double)hl.x);
      sm[i] = hl;
    }
    mpair = sm;
  }
  if (threadIdx.x == 0) s_maxbits = 0u;
  __syncthreads();

  const long long nvec = p.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  bool bad = false;

  // ---- reference exponent of this CTA: from its first tile ---------------------------------------
  unsigned lo_bits = 0, span_bits = 0; double scale = 1.0, inv_scale = 1.0;
  if (IS_SUM) {
    unsigned m = 0;
    if (v < nvec) {
      int4 av = ldg_stream16((const int4*)p.a + v);
      unsigned b[4] = {(unsigned)av.x & 0x7fffffffu, (unsigned)av.y & 0x7fffffffu, (unsigned)av.z & 0x7fffffffu, (unsigned)av.w & 0x7fffffffu};
      long long gs[4] = {0, 0, 0, 0};
      if (xmeans) LabelVec<LT>::load(p.labels, v, gs);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned bb = b[k];
        if (p.square) { float f = __uint_as_float(bb); bb = __float_as_uint(f * f); }
        if (xmeans) {
          const float xk = __int_as_float(k == 0 ? av.x : (k == 1 ? av.y : (k == 2 ? av.z : av.w)));
          const double d = (gs[k] >= 0 && gs[k] < p.size) ? (double)xk - xmeans[gs[k]] : 0.0;
          bb = __float_as_uint((float)(d * d));
        }
        if (bb < 0x7f800000u && bb > m) m = bb;
      }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&s_maxbits, m);
    __syncthreads();
    int eb = (int)(s_maxbits >> 23);            // biased exponent of the largest finite |value| seen
    if (eb == 0) eb = 127;                      // tile was all zero / non-finite: assume O(1) data
    eb += 1;                                    // values with biased exponent < eb convert
    // squared deviations (second pass of var / std, f32 results) spread over twice as many binades as the data, and on a
    // small table their bypass REDs would queue on a few hot addresses (normal data, 100 groups: 1.6 % of the terms, 7 ms).
    // Exactness is not needed here: a term as small as 2^-12 (one word) / 2^-22 (two words) of the reference is rounded
    // to the grid -- relative error of that term <= 2^-18, of the group's sum far less; only what is smaller still takes
    // the bypass
    const int window = xmeans ? (ONEWORD ? 12 : 22) : FXW;
    int lo_e = eb - window; if (lo_e < 1) lo_e = 1;
    lo_bits = (unsigned)lo_e << 23;
    span_bits = (unsigned)(eb - lo_e) << 23;
    // q = v * 2^(FXF - (eb-127)) ; exact for in-window values
    scale = __longlong_as_double((long long)(1023 + FXF - (eb - 127)) << 52);
    inv_scale = __longlong_as_double((long long)(1023 - FXF + (eb - 127)) << 52);
  }
  const double hi_unit = inv_scale * 4294967296.0;        // FM_SUM1: what one carry out of the 32-bit word is worth

  // accumulate straight into the global tables (uncovered groups, ragged tail)
  const unsigned long long pol = PARTIAL ? l2_evict_last_policy() : 0ull;
  auto direct_global = [&](long long g, float x) {
    const bool nan = x != x;
    if (LOAD_A && nan && nanskip) return;
    if (MODE == FM_LEN) { if (PARTIAL) red_add_u64_hint((unsigned long long*)(p.t.t1 + g), 1ull, pol); else atomicAdd((unsigned long long*)(p.t.t1 + g), 1ull); }
    else if (MODE == FM_MINMAX) {
      const long long kmax = nan ? AGC_EMPTY_MIN : key_of((double)x), kmin = nan ? AGC_EMPTY_MAX : key_of((double)x);
      atomicMax(p.t.t0i + g, kmax); atomicMin((long long*)p.t.t2 + g, kmin);
    } else if (MODE == FM_MAX || MODE == FM_MIN) {
      const long long k64 = nan ? (MODE == FM_MAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX) : key_of((double)x);
      // (test-then-update -- an L2 read of the current extreme before the RED -- was measured HERE and lost: max at 10^7 groups
      //  6.8 -> 20.7 ms.  The read must return before the thread knows whether to send the RED, eight dependent round trips
      //  per thread iteration; a RED is fire-and-forget.  The general kernel keeps it: there one element per thread is in flight.)
      if (PARTIAL) { if (MODE == FM_MAX) red_max_s64_hint(p.t.t0i + g, k64, pol); else red_min_s64_hint(p.t.t0i + g, k64, pol); }
      else { if (MODE == FM_MAX) atomicMax(p.t.t0i + g, k64); else atomicMin(p.t.t0i + g, k64); }
    } else {
      if (p.square) x = x * x;
      if ((__float_as_uint(x) & 0x7fffffffu) != 0u) { if (PARTIAL) red_add_f64_hint(p.t.t0f + g, (double)x, pol); else atomicAdd(p.t.t0f + g, (double)x); }
      if (HAS_CNT) { if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + g), 1ull); if (p.want_b1) p.t.b1[g] = 1; }
    }
  };
  auto update = [&](long long g, float x) {
    if (g < 0 || g >= p.size) { bad = true; return; }
    if (IS_SUM && xmeans) {
      // nan-forms drop ORIGINAL NaNs only: a NaN deviation of a non-NaN element (+-inf in the group) must propagate, as in
      // the reference's (a - mean)**2 (aggregate_numpy.py:184-186)
      if (nanskip && x != x) return;
      if (mpair) { const float2 hl = mpair[g]; const float d = (x - hl.x) - hl.y; x = d * d; }
      else { const double d = (double)x - __ldg(xmeans + g); x = (float)(d * d); }
    }
    if (PARTIAL && g >= p.cov) { direct_global(g, x); return; }
    const int gi = ((int)g << rs) | rsel;
    if (MODE == FM_LEN) {
      if (LOAD_A && nanskip && x != x) return;
      atomicAdd(w0 + gi, 1u);
    } else if (MODE == FM_MINMAX) {
      int kx, kn;
      if (x != x) { if (nanskip) return; kx = 0x7fffffff; kn = (int)0x80000000; }           // NaN wins both
      else kx = kn = key32_of(x);
      atomicMax((int*)w0 + gi, kx); atomicMin((int*)w1 + gi, kn);
    } else if (MODE == FM_MAX || MODE == FM_MIN) {
      int k;
      if (x != x) { if (nanskip) return; k = MODE == FM_MAX ? 0x7fffffff : (int)0x80000000; }   // NaN wins
      else k = key32_of(x);
      if (MODE == FM_MAX) atomicMax((int*)w0 + gi, k); else atomicMin((int*)w0 + gi, k);
    } else {
      if (nanskip && !xmeans && x != x) return;
      if (p.square) x = x * x;
      unsigned ab = __float_as_uint(x) & 0x7fffffffu;
      if (ab - lo_bits < span_bits) {
        if (ONEWORD) {
          const int q = __double2int_rn((double)x * scale);                 // |q| < 2^29, exact
          const unsigned old = atomicAdd(w0 + gi, (unsigned)q);
          const int net = (q >> 31) + (int)(old + (unsigned)q < old);         // carry (+1) / borrow (-1) out of the word
          if (net) atomicAdd(p.t.t0f + g, net > 0 ? hi_unit : -hi_unit);
        } else {
          long long q = __double2ll_rn((double)x * scale);
          fx_add(w0 + gi, w1 + gi, q);
        }
      } else if (ab != 0u) {                        // out of window (huge / tiny / inf / NaN): f64, on chip if there is room
        // (the side table's address is recomputed from the kernel parameters and the add is an out-of-line call: the hot
        // loop of the 32-register variant has no room for another live pointer)
        if (xside) smem_side_add(sw + ((WORDS * (p.cov << p.rshift) + 1) & ~1), gi, (double)x);
        else atomicAdd(p.t.t0f + g, (double)x);
      }
      if (HAS_CNT) atomicAdd(wc + gi, 1u);
    }
  };

  for (; v < nvec; v += stride * U) {
    long long g[U][4]; float x[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long vu = v + u * stride;
#pragma unroll
      for (int k = 0; k < 4; ++k) { g[u][k] = 0; x[u][k] = 0.f; }
      if (u == 0 || vu < nvec) {
        LabelVec<LT>::load(p.labels, vu, g[u]);
        if (LOAD_A) {
          int4 av = ldg_stream16((const int4*)p.a + vu);
          x[u][0] = __int_as_float(av.x); x[u][1] = __int_as_float(av.y); x[u][2] = __int_as_float(av.z); x[u][3] = __int_as_float(av.w);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (u > 0 && v + u * stride >= nvec) break;
#pragma unroll
      for (int k = 0; k < 4; ++k) update(g[u][k], x[u][k]);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {          // tail elements
    long long i = (nvec << 2) + threadIdx.x;
    update(LabelVec<LT>::one(p.labels, i), LOAD_A ? p.a[i] : 0.f);
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.status[AGC_ST_REGIME] = MODE == FM_SUMCNT1 ? 19 : (MODE == FM_MINMAX ? 30 : 10 + MODE);
  __syncthreads();

  // ---- merge the CTA's table into the global accumulator tables -----------------------------------
  // one thread per group: combine its replicas in shared memory, then ONE RED per touched group
  const int R = 1 << rs;
  for (int gl = threadIdx.x; gl < p.cov; gl += blockDim.x) {
    const int e0 = gl << rs;
    const int gi = gl;
    if (MODE == FM_LEN) {
      unsigned long long c = 0;
      for (int r = 0; r < R; ++r) c += w0[e0 + r];
      if (c) atomicAdd((unsigned long long*)(p.t.t1 + gi), c);
    } else if (MODE == FM_MINMAX) {
      int kx = (int)0x80000000, kn = 0x7fffffff;
      for (int r = 0; r < R; ++r) { kx = max(kx, (int)w0[e0 + r]); kn = min(kn, (int)w1[e0 + r]); }
      if (kx != (int)0x80000000) atomicMax(p.t.t0i + gi, kx == 0x7fffffff ? AGC_EMPTY_MIN : key_of((double)unkey_f32(kx)));
      if (kn != 0x7fffffff) atomicMin((long long*)p.t.t2 + gi, kn == (int)0x80000000 ? AGC_EMPTY_MAX : key_of((double)unkey_f32(kn)));
    } else if (MODE == FM_MAX || MODE == FM_MIN) {
      int k = (int)init0;
      for (int r = 0; r < R; ++r) { int kr = (int)w0[e0 + r]; k = MODE == FM_MAX ? max(k, kr) : min(k, kr); }
      if (k != (int)init0) {
        long long k64;
        if (MODE == FM_MAX) k64 = k == 0x7fffffff ? AGC_EMPTY_MIN : key_of((double)unkey_f32(k));
        else k64 = k == (int)0x80000000 ? AGC_EMPTY_MAX : key_of((double)unkey_f32(k));
        if (MODE == FM_MAX) atomicMax(p.t.t0i + gi, k64); else atomicMin(p.t.t0i + gi, k64);
      }
    } else {
      long long q = 0; unsigned long long c = 0;
      for (int r = 0; r < R; ++r) {
        if (ONEWORD) q += (long long)w0[e0 + r] - 0x80000000ll;                      // residual around the 2^31 start value
        else q += (long long)(((unsigned long long)w1[e0 + r] << 32) | w0[e0 + r]);  // <= 32 * 2^62 / 32: no overflow (FX_F headroom)
        if (HAS_CNT) c += wc[e0 + r];
      }
      double tot = (double)q * inv_scale;
      bool any = q != 0;
      if (side) { for (int r = 0; r < R; ++r) { const double sv = side[e0 + r]; if (sv != 0.0) { tot += sv; any = true; } } }
      if (any) atomicAdd(p.t.t0f + gi, tot);
      if (HAS_CNT && c) {
        if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + gi), c);
        if (p.want_b1) p.t.b1[gi] = 1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_len16: counts with 16-bit counters, two per shared-memory word -- ~100 K groups in 195 KB.
// `len`, and the count pass of the split mean / var on tables the one-word sum kernel also covers.
// An element reads its counter word (plain LDS), then adds 1 to its half with a shared-memory RED (no return value,
// no barrier anywhere).  A thread that READS a half >= 0x8000 takes 0x8000 out of it with a CAS and forwards 32768 to
// the global count before it adds.  Why a half can never carry into its neighbour: an add is "blind" only if its LDS
// came before the half crossed 0x8000; a thread issues the LDS of a batch of 8 elements before their adds, so at most
// 1024 x 8 blind adds can follow a crossing -- the half stays below 0x8000 + 8192 + 8192 < 0x10000 until the next
// reader repairs it, and the final merge reads all 16 bits.
// ---------------------------------------------------------------------------------------------
template <class LT, bool LOAD_A>
__global__ void __launch_bounds__(FAST_THREADS, 1) k_len16(const FastParams p) {
  if (p.choose && *p.choose != p.want) return;
  extern __shared__ unsigned sw[];
  const int W = (p.cov + 1) >> 1;                              // groups [0, cov) have a counter, the rest go straight to the global table
  for (int i = threadIdx.x; i < W; i += blockDim.x) sw[i] = 0u;
  __syncthreads();
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sw);
  const long long nvec = p.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  const unsigned long long pol = l2_evict_last_policy();
  bool bad = false;
  // take 0x8000 out of the half `sh` of word `wi` (if it still is >= 0x8000) and forward it to the global count
  auto repair = [&](int wi, int sh, long long g) {
    unsigned o = *(volatile unsigned*)(sw + wi);
    while (((o >> sh) & 0xffffu) >= 0x8000u) {
      const unsigned seen = atomicCAS(sw + wi, o, o - (0x8000u << sh));
      if (seen == o) { red_add_u64_hint((unsigned long long*)(p.t.t1 + g), 0x8000ull, pol); break; }
      o = seen;
    }
  };
  auto update = [&](long long g, float x) {                  // one element at a time (ragged tail)
    if ((unsigned long long)g >= (unsigned long long)p.size) { bad = true; return; }
    if (LOAD_A && nanskip && x != x) return;
    if (g >= p.cov) { red_add_u64_hint((unsigned long long*)(p.t.t1 + g), 1ull, pol); return; }
    const int sh = ((int)g & 1) << 4, wi = (int)g >> 1;
    if (((*(volatile unsigned*)(sw + wi) >> sh) & 0xffffu) >= 0x8000u) repair(wi, sh, g);
    reds_add_u32(sbase + wi * 4, 1u << sh);
  };
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += 2 * stride) {
    long long g[2][4]; float x[2][4];
    const bool two = v + stride < nvec;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { g[u][k] = -1; x[u][k] = 0.f; }
      if (u == 0 || two) {
        LabelVec<LT>::load(p.labels, v + u * stride, g[u]);
        if (LOAD_A) {
          int4 av = ldg_stream16((const int4*)p.a + v + u * stride);
          x[u][0] = __int_as_float(av.x); x[u][1] = __int_as_float(av.y); x[u][2] = __int_as_float(av.z); x[u][3] = __int_as_float(av.w);
        }
      }
    }
    unsigned cur[8];                                          // counter words of the batch, all read before the first add
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const long long gg = g[e >> 2][e & 3];
      const bool ok = (unsigned long long)gg < (unsigned long long)p.size;
      if (!ok && (e < 4 || two)) bad = true;
      const bool take = ok && !(LOAD_A && nanskip && x[e >> 2][e & 3] != x[e >> 2][e & 3]);
      // 0xffffffff: skip; covered groups read their word; uncovered groups are marked with 0xfffffffe
      cur[e] = !take ? 0xffffffffu : (gg < p.cov ? *(volatile unsigned*)(sw + ((int)gg >> 1)) : 0xfffffffeu);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (cur[e] == 0xffffffffu) continue;
      const long long gg = g[e >> 2][e & 3];
      if (gg >= p.cov) { red_add_u64_hint((unsigned long long*)(p.t.t1 + gg), 1ull, pol); continue; }
      const int sh = ((int)gg & 1) << 4, wi = (int)gg >> 1;
      if (((cur[e] >> sh) & 0xffffu) >= 0x8000u) repair(wi, sh, gg);
      reds_add_u32(sbase + wi * 4, 1u << sh);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {
    const long long i = (nvec << 2) + threadIdx.x;
    update(LabelVec<LT>::one(p.labels, i), LOAD_A ? p.a[i] : 0.f);
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.status[AGC_ST_REGIME] = 16;
  __syncthreads();
  for (int g = threadIdx.x; g < p.cov; g += blockDim.x) {
    const unsigned c = (sw[g >> 1] >> ((g & 1) << 4)) & 0xffffu;
    if (c) atomicAdd((unsigned long long*)(p.t.t1 + g), (unsigned long long)c);
  }
}
template <class LT, bool LOAD_A>
inline int launch_len16(const FastParams& p, cudaStream_t st, int sms) {
  auto kern = k_len16<LT, LOAD_A>;
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dyn_smem(kern, 227 * 1024 - 64, attr_done)) return -30;
  const size_t smem = (size_t)((p.cov + 1) / 2) * 4;
  AGC_COUNT_LAUNCH(1), kern<<<sms, FAST_THREADS, smem, st>>>(p);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// k_maxfilter: max / min for tables of up to 113 K groups -- a 16-bit FILTER per group in shared memory.
// top[g] holds the upper 16 bits of the order-preserving key of some element of group g this CTA has already sent to
// the global table.  An element whose upper 16 bits are smaller cannot be the maximum and is dropped after one LDS;
// otherwise it is sent (exact 64-bit key, RED.MAX) and, if larger, recorded with a plain 16-bit store.  No shared
// atomics: a lost store only makes the filter weaker, never wrong, because every stored value comes from an element
// that did reach the global table.  A CTA sees ~70 elements of a group at 10^5 groups and N = 2^30, so about 7 % of
// the elements still pass (record highs + ties in the upper 16 bits); the rest never leave the SM, for uniform,
// skewed and sorted labels alike.  NaN "wins" in the reference: it maps to the top code 0xFFFF, after which the
// group's filter rejects everything.  min uses the complemented key.  (ncu: half of the stall samples wait for the
// global loads -- one CTA of 1024 threads per SM; a register double-buffered prefetch spilled and was slower.)
// ---------------------------------------------------------------------------------------------
template <bool ISMAX, class LT, bool TIE>
__global__ void __launch_bounds__(FAST_THREADS, 1) k_maxfilter(const FastParams p) {
  if (p.choose && *p.choose != p.want) return;              // 0: uniform labels (TIE = false), 1: skewed (TIE = true); runs go to k_runs
  extern __shared__ unsigned sw[];
  unsigned short* top = (unsigned short*)sw;
  const int W = (p.cov + 1) >> 1;                              // groups [0, cov) have a filter word, the rest always pass
  for (int i = threadIdx.x; i < W; i += blockDim.x) sw[i] = 0u;
  __syncthreads();
  const long long nvec = p.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  const unsigned long long pol = l2_evict_last_policy();
  bool bad = false;
  auto update = [&](long long g, float x) {
    if ((unsigned long long)g >= (unsigned long long)p.size) { bad = true; return; }
    const bool nan = x != x;
    if (nan && nanskip) return;
    const int k32 = nan ? (ISMAX ? 0x7fffffff : (int)0x80000000) : key32_of(x);              // NaN wins
    const unsigned u = ISMAX ? ((unsigned)k32 ^ 0x80000000u) : ~((unsigned)k32 ^ 0x80000000u);  // larger u = better
    const unsigned t = u >> 16;
    const unsigned cur = g < p.cov ? (unsigned)top[g] : 0u;
    if (t < cur || cur == 0xffffu) return;
    const long long k64 = nan ? (ISMAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX) : key_of((double)x);
    if (ISMAX) red_max_s64_hint(p.t.t0i + g, k64, pol); else red_min_s64_hint(p.t.t0i + g, k64, pol);
    if (t > cur && g < p.cov) top[g] = (unsigned short)t;
  };
  // A batch of MF_B elements reads its filter words first (independent LDS, one latency) and decides afterwards: a
  // filter word that an earlier element of the same batch has just raised is read stale, which only lets a few more
  // elements through.  (One element at a time, the chain LDS -> compare -> STS of 8 elements was a third of the
  // kernel's stall samples.)
  constexpr int MF_B = 8;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += 2 * stride) {
    long long g[2][4]; int4 av[2];
    const bool two = v + stride < nvec;
    LabelVec<LT>::load(p.labels, v, g[0]);
    av[0] = ldg_stream16((const int4*)p.a + v);
    if (two) { LabelVec<LT>::load(p.labels, v + stride, g[1]); av[1] = ldg_stream16((const int4*)p.a + v + stride); }
    else {
#pragma unroll
      for (int k = 0; k < 4; ++k) g[1][k] = -1;
      av[1] = make_int4(0, 0, 0, 0);
    }
#pragma unroll
    for (int b0 = 0; b0 < 8; b0 += MF_B) {
      unsigned t[MF_B], cur[MF_B];
#pragma unroll
      for (int e = 0; e < MF_B; ++e) {
        const int u = (b0 + e) >> 2, k = (b0 + e) & 3;
        const long long gg = g[u][k];
        const float x = __int_as_float(k == 0 ? av[u].x : (k == 1 ? av[u].y : (k == 2 ? av[u].z : av[u].w)));
        const bool nan = x != x;
        const int k32 = nan ? (ISMAX ? 0x7fffffff : (int)0x80000000) : key32_of(x);            // NaN wins
        const unsigned uu = ISMAX ? ((unsigned)k32 ^ 0x80000000u) : ~((unsigned)k32 ^ 0x80000000u);  // larger uu = better
        t[e] = uu >> 16;
        const bool ok = (unsigned long long)gg < (unsigned long long)p.size;
        if (!ok && (u == 0 || two)) bad = true;
        cur[e] = (ok && !(nan && nanskip)) ? (gg < p.cov ? (unsigned)top[gg] : 0u) : 0xffffu;  // 0xffff rejects everything
      }
#pragma unroll
      for (int e = 0; e < MF_B; ++e) {
        if (t[e] < cur[e] || cur[e] == 0xffffu) continue;
        const int u = (b0 + e) >> 2, k = (b0 + e) & 3;
        const long long gg = g[u][k];
        const float x = __int_as_float(k == 0 ? av[u].x : (k == 1 ? av[u].y : (k == 2 ? av[u].z : av[u].w)));
        const long long k64 = x != x ? (ISMAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX) : key_of((double)x);
        if (TIE && t[e] == cur[e]) {
          // skewed labels (the probe says so): a tie in the upper 16 bits looks at the exact value in the global table
          // first (L2 read; a stale value only costs a redundant RED).  Keeps the ties of a HOT group -- thousands per
          // CTA -- from queueing on one L2 address (Zipf, 10^5 groups: 277-325 -> 373 Gelem/s); on uniform labels the
          // extra round trip per tie costs 15 %, so that variant runs without it.
          const long long seen = __ldcg(p.t.t0i + gg);
          if (ISMAX ? k64 <= seen : k64 >= seen) continue;
        }
        if (ISMAX) red_max_s64_hint(p.t.t0i + gg, k64, pol); else red_min_s64_hint(p.t.t0i + gg, k64, pol);
        if (t[e] > cur[e] && gg < p.cov) top[gg] = (unsigned short)t[e];
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {
    const long long i = (nvec << 2) + threadIdx.x;
    update(LabelVec<LT>::one(p.labels, i), p.a[i]);
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.status[AGC_ST_REGIME] = ISMAX ? 17 : 18;
}
template <bool ISMAX, class LT, bool TIE>
inline int launch_maxfilter(const FastParams& p, cudaStream_t st, int sms) {
  auto kern = k_maxfilter<ISMAX, LT, TIE>;
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dyn_smem(kern, 227 * 1024 - 64, attr_done)) return -30;
  const size_t smem = (size_t)((p.cov + 1) / 2) * 4;
  AGC_COUNT_LAUNCH(1), kern<<<sms, FAST_THREADS, smem, st>>>(p);
  return 1;
}

template <int MODE, class LT, bool LOAD_A>
__global__ void __launch_bounds__(FAST_THREADS, 2) k_cache(const FastParams p) {
  if (p.choose && *p.choose != p.want) return;
  if (p.choose2 && *p.choose2 != p.want2) return;
  extern __shared__ unsigned sw[];
  __shared__ unsigned s_maxbits;
  constexpr int WORDS = MODE == FM_SUM ? 2 : (MODE == FM_SUMCNT ? 3 : 1);
  const unsigned S = (unsigned)p.nslots;
  unsigned* tag = sw;             // 0 = free, else group + 1
  unsigned* w0 = sw + S;
  unsigned* w1 = sw + 2 * S;
  unsigned* w2 = sw + 3 * S;
  // sums: an f64 side accumulator per slot for the values of a cached key that fall outside the fixed-point window -- as
  // global REDs the out-of-window values of a HOT key queue on one L2 address (Zipf labels, data over 20 binades: 14x slower)
  // (only when the probe saw values spread over more binades than the window: the slots it costs lower the hit rate)
  const bool HAS_SIDE = (MODE == FM_SUM || MODE == FM_SUMCNT) && p.side;
  unsigned* side_words = sw + (1 + WORDS) * S;               // S is even: 8-byte aligned
  const unsigned init0 = MODE == FM_MAX ? 0x80000000u : (MODE == FM_MIN ? 0x7fffffffu : 0u);
  for (unsigned i = threadIdx.x; i < (1 + WORDS + (HAS_SIDE ? 2 : 0)) * S; i += blockDim.x) sw[i] = (i >= S && i < 2 * S) ? init0 : 0u;
  if (threadIdx.x == 0) s_maxbits = 0u;
  __syncthreads();

  const long long nvec = p.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  bool bad = false;

  unsigned lo_bits = 0, span_bits = 0; double scale = 1.0, inv_scale = 1.0;
  if (MODE == FM_SUM || MODE == FM_SUMCNT) {
    unsigned m = 0;
    if (v < nvec) {
      int4 av = ldg_stream16((const int4*)p.a + v);
      unsigned b[4] = {(unsigned)av.x & 0x7fffffffu, (unsigned)av.y & 0x7fffffffu, (unsigned)av.z & 0x7fffffffu, (unsigned)av.w & 0x7fffffffu};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned bb = b[k];
        if (p.square) { float f = __uint_as_float(bb); bb = __float_as_uint(f * f); }
        if (bb < 0x7f800000u && bb > m) m = bb;
      }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&s_maxbits, m);
    __syncthreads();
    int eb = (int)(s_maxbits >> 23);
    if (eb == 0) eb = 127;
    eb += 1;
    int lo_e = eb - FX_WINDOW; if (lo_e < 1) lo_e = 1;
    lo_bits = (unsigned)lo_e << 23;
    span_bits = (unsigned)(eb - lo_e) << 23;
    scale = __longlong_as_double((long long)(1023 + FX_F - (eb - 127)) << 52);
    inv_scale = __longlong_as_double((long long)(1023 - FX_F + (eb - 127)) << 52);
  }

  const unsigned long long pol = l2_evict_last_policy();       // misses: keep the global table in L2 while the input streams by
  auto update = [&](long long g, float x) {
    if (g < 0 || g >= p.size) { bad = true; return; }
    if (LOAD_A && x != x && nanskip) return;
    if (MODE == FM_SUM || MODE == FM_SUMCNT) { if (p.square) x = x * x; }
    // ---- cache lookup ----
    // 2-way: the key may live in slot s or in its neighbour s ^ 1 (two hot keys that hash together no longer fight:
    // a permanent loser would send ALL its elements to one global address)
    const unsigned want = (unsigned)g + 1u;
    unsigned s = __umulhi((unsigned)g * 0x9E3779B1u, S);
    unsigned t = tag[s];
    if (t == 0u) { t = atomicCAS(tag + s, 0u, want); if (t == 0u) t = want; }
    if (t != want) {
      const unsigned s2 = s ^ 1u;
      unsigned t2 = tag[s2];
      if (t2 == 0u) { t2 = atomicCAS(tag + s2, 0u, want); if (t2 == 0u) t2 = want; }
      if (t2 == want) { s = s2; t = t2; }
    }
    const bool hit = t == want;
    if (MODE == FM_LEN) {
      if (hit) atomicAdd(w0 + s, 1u); else red_add_u64_hint((unsigned long long*)(p.t.t1 + g), 1ull, pol);
    } else if (MODE == FM_MAX || MODE == FM_MIN) {
      const bool nan = x != x;
      if (hit) {
        int k = nan ? (MODE == FM_MAX ? 0x7fffffff : (int)0x80000000) : key32_of(x);
        if (MODE == FM_MAX) atomicMax((int*)w0 + s, k); else atomicMin((int*)w0 + s, k);
      } else {
        long long k64 = nan ? (MODE == FM_MAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX) : key_of((double)x);
        if (MODE == FM_MAX) red_max_s64_hint(p.t.t0i + g, k64, pol); else red_min_s64_hint(p.t.t0i + g, k64, pol);
      }
    } else {
      const unsigned ab = __float_as_uint(x) & 0x7fffffffu;
      if (hit && (ab - lo_bits < span_bits)) {
        fx_add(w0 + s, w1 + s, __double2ll_rn((double)x * scale));
      } else if (ab != 0u) {
        if (hit && HAS_SIDE) smem_side_add(side_words, (int)s, (double)x);   // cached key, outside the exact window: stays on chip
        else red_add_f64_hint(p.t.t0f + g, (double)x, pol);                   // miss (or narrow data: no side table)
      }
      if (MODE == FM_SUMCNT) {
        if (hit) atomicAdd(w2 + s, 1u);
        else { if (p.want_cnt) red_add_u64_hint((unsigned long long*)(p.t.t1 + g), 1ull, pol); if (p.want_b1) p.t.b1[g] = 1; }
      }
    }
  };

  for (; v < nvec; v += stride) {
    long long g[4];
    LabelVec<LT>::load(p.labels, v, g);
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (LOAD_A) {
      int4 av = ldg_stream16((const int4*)p.a + v);
      x[0] = __int_as_float(av.x); x[1] = __int_as_float(av.y); x[2] = __int_as_float(av.z); x[3] = __int_as_float(av.w);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) update(g[k], x[k]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {
    long long i = (nvec << 2) + threadIdx.x;
    update(LabelVec<LT>::one(p.labels, i), LOAD_A ? p.a[i] : 0.f);
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.status[AGC_ST_REGIME] = 20 + MODE;
  __syncthreads();

  for (unsigned s = threadIdx.x; s < S; s += blockDim.x) {
    const unsigned t = tag[s];
    if (t == 0u) continue;
    const long long gi = (long long)t - 1;
    if (MODE == FM_LEN) {
      unsigned c = w0[s];
      if (c) atomicAdd((unsigned long long*)(p.t.t1 + gi), (unsigned long long)c);
    } else if (MODE == FM_MAX || MODE == FM_MIN) {
      int k = (int)w0[s];
      if (k != (int)init0) {
        long long k64;
        if (MODE == FM_MAX) k64 = k == 0x7fffffff ? AGC_EMPTY_MIN : key_of((double)unkey_f32(k));
        else k64 = k == (int)0x80000000 ? AGC_EMPTY_MAX : key_of((double)unkey_f32(k));
        if (MODE == FM_MAX) atomicMax(p.t.t0i + gi, k64); else atomicMin(p.t.t0i + gi, k64);
      }
    } else {
      long long q = (long long)(((unsigned long long)w1[s] << 32) | w0[s]);
      const double sv = HAS_SIDE ? ((const double*)side_words)[s] : 0.0;
      if (q != 0 || sv != 0.0) atomicAdd(p.t.t0f + gi, (double)q * inv_scale + sv);
      if (MODE == FM_SUMCNT) {
        unsigned c = w2[s];
        if (c) {
          if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + gi), (unsigned long long)c);
          if (p.want_b1) p.t.b1[gi] = 1;
        }
      }
    }
  }
}

template <int MODE, class LT, bool LOAD_A>
inline int launch_cache(const FastParams& p, int ctas_per_sm, size_t smem, cudaStream_t st, int sms) {
  auto kern = k_cache<MODE, LT, LOAD_A>;
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dyn_smem(kern, 227 * 1024 - 64, attr_done)) return -30;
  AGC_COUNT_LAUNCH(1), kern<<<sms * ctas_per_sm, FAST_THREADS, smem, st>>>(p);
  return 1;
}

template <int MODE, class LT, bool LOAD_A, int U, bool PARTIAL, bool X>
inline int launch_fast_x(const FastParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kern = k_fast<MODE, LT, LOAD_A, U, PARTIAL, X>;
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dyn_smem(kern, 227 * 1024 - 64, attr_done)) return -30;
  AGC_COUNT_LAUNCH(1), kern<<<grid, FAST_THREADS, smem, st>>>(p);
  return 1;
}
template <int MODE, class LT, bool LOAD_A, int U, bool PARTIAL>
inline int launch_fast_u(const FastParams& p, int grid, size_t smem, cudaStream_t st) {
  constexpr bool SUMMODE = MODE == FM_SUM || MODE == FM_SUMCNT || MODE == FM_SUM1 || MODE == FM_SUMCNT1;
  if constexpr (SUMMODE && LOAD_A && !PARTIAL) {
    if (p.side || p.means) return launch_fast_x<MODE, LT, LOAD_A, U, PARTIAL, true>(p, grid, smem, st);
  }
  return launch_fast_x<MODE, LT, LOAD_A, U, PARTIAL, false>(p, grid, smem, st);
}
template <int MODE, class LT, bool LOAD_A>
inline int launch_fast(const FastParams& p, int ctas_per_sm, size_t smem, cudaStream_t st, int sms) {
  const int grid = sms * ctas_per_sm;
  if (p.cov < p.size) return launch_fast_u<MODE, LT, LOAD_A, 2, true>(p, grid, smem, st);      // always 1 CTA/SM
  return ctas_per_sm == 1 ? launch_fast_u<MODE, LT, LOAD_A, 2, false>(p, grid, smem, st) : launch_fast_u<MODE, LT, LOAD_A, 1, false>(p, grid, smem, st);
}

}  // namespace agc
