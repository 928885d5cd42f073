// agc_dispatch.cuh -- which accumulate kernel runs: the device-side label probe, try_fast_path() and try_fast_pass2().
//
// Form 1, f32 values, int32 / int64 labels, sum / sumofsquares / mean / var / min / max / len (+ nan-forms):
//   table <= 195 KB of shared memory per SM ...... k_fast (agc_fast.cuh): whole table per CTA, replicated when small, two
//       (AGC_SMEM1 / AGC_SMEM2: more than 196 KB     CTAs per SM up to 97 KB each; sums of <= 12.4 K groups carry an f64 side
//        per SM costs a third of the load bandwidth)  accumulator; mean / var on 16.6 K - 24.9 K groups: one-word sum + count
//   larger tables (no host round trip: the probe writes status[4] -- 0 uniform, 1 skewed, 2 runs -- and status[7] -- values
//   spread beyond the fixed-point window --, every candidate kernel is launched and the ones not chosen return at once)
//       labels in runs (sorted / contiguous) ..... k_runs  (agc_runs.cuh): segmented warp reduce
//       skewed labels ............................ k_cache (agc_fast.cuh): per-CTA hot-key cache + global RED (max / min up to
//                                                  113 K groups: k_maxfilter over the whole table, with the exact test on ties)
//       otherwise: sum ........................... k_fast one-word carry-forwarding sums, partial table + global RED
//                  max / min ..................... k_maxfilter: 16-bit per-group filter + exact RED for what passes
//                  len ........................... k_len16: 16-bit counters, read-check-repair overflow forwarding
//                  mean / var pass 1 ............. one-word sum pass + count pass (k_len16)
//       partial tables take 195 KB while that covers >= 90 % of the groups, else 226 KB (pick_cov)
//   second pass of var / std, f32 result ......... k_fast on (float)((a - mean)^2) up to 49.9 K groups (try_fast_pass2)
// Form 3 on the last axis .......................... k_form3 (agc_form3.cuh)
// everything else .................................. k_accumulate_small (<= 16.5 K groups) / k_accumulate (agc_small / agc_generic)
#pragma once
#include "agc_fast.cuh"
#include "agc_form3.cuh"
#include "agc_runs.cuh"
#include "agc_team.cuh"
#include <mutex>

namespace agc {

// ---------------------------------------------------------------------------------------------
// label-distribution probe: 32 Ki strided samples -> (a) largest share of one hash bucket,
// (b) share of positions whose right neighbour carries the same label.  status[4] = 1 when the
// labels are skewed or come in runs (the per-CTA hot-key cache wins there), else 0.
// ---------------------------------------------------------------------------------------------
constexpr int PROBE_SAMPLES = 32768;                // at n >= 2^23; fewer for smaller inputs (n / 256, at least 4096)
// A hash bucket holds 8 samples on average for uniform labels (max over 4096 buckets ~ 22).  One group with >= 0.1 % of
// the elements adds >= 33: beyond that share its uncovered elements serialise on ONE L2 address (measured 1.5 ns per
// RED: a 0.5 % group doubles the partial-table kernel's time, 1.5 % makes it 5x) and the hot-key cache wins.
constexpr int PROBE_SKEW_COUNT = 40;
// The probe is spread over samples / 1024 CTAs (one sample per thread: a single dependent round trip to memory instead of
// the 32 a lone CTA needed -- 77 us per call in round 1, 13 % of a 2^27-element step).  Every CTA adds its bucket and
// exponent histograms to a scratch area in the caller's workspace; the last one to finish (a ticket counter) reads the
// totals and writes status[4..7] (the dispatcher zeroes the scratch area before every probe).
constexpr int PROBE_SCRATCH_WORDS = 4096 + 256 + 8;  // bucket histogram, exponent histogram, {adjacent-equal count, ticket}
template <class LT>
__global__ void __launch_bounds__(1024) k_skew_probe(const void* labels, long long n, int* status, int samples, int thresh,
                                                     unsigned* scratch, const float* a = nullptr) {
  __shared__ unsigned cnt[4096];
  __shared__ unsigned ecnt[256];
  __shared__ unsigned s_adj, s_last;
  for (int i = threadIdx.x; i < 4096; i += 1024) cnt[i] = 0;
  if (threadIdx.x < 256) ecnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_adj = 0; s_last = 0; }
  __syncthreads();
  const long long stride = n / samples;                // >= 2 (callers guarantee n >= 65536 and samples <= n / 16)
  const long long i = ((long long)blockIdx.x * 1024 + threadIdx.x) * stride;
  const long long g = LabelVec<LT>::one(labels, i), gn = LabelVec<LT>::one(labels, i + 1);
  atomicAdd(&cnt[((unsigned)g * 0x9E3779B1u) >> 20], 1u);
  unsigned adj = g == gn;
  if (a) {                                              // biased exponent of the sampled value (zero / inf / NaN aside)
    const unsigned bits = __float_as_uint(a[i]) & 0x7fffffffu;
    const unsigned e = bits >> 23;
    if (bits != 0u && e != 255u) atomicAdd(&ecnt[e], 1u);
  }
  adj = __reduce_add_sync(0xffffffffu, adj);
  if ((threadIdx.x & 31) == 0 && adj) atomicAdd(&s_adj, adj);
  __syncthreads();
  unsigned* g_cnt = scratch; unsigned* g_ecnt = scratch + 4096; unsigned* g_misc = scratch + 4096 + 256;
  for (int k = threadIdx.x; k < 4096; k += 1024) if (cnt[k]) atomicAdd(g_cnt + k, cnt[k]);
  if (threadIdx.x < 256 && ecnt[threadIdx.x]) atomicAdd(g_ecnt + threadIdx.x, ecnt[threadIdx.x]);
  if (threadIdx.x == 0 && s_adj) atomicAdd(g_misc, s_adj);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(g_misc + 1, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  // ---- the last CTA: totals -> decision, scratch back to zero ----
  __threadfence();
  __shared__ unsigned s_max, s_emax, s_low;
  if (threadIdx.x == 0) { s_max = 0; s_emax = 0; s_low = 0; }
  __syncthreads();
  unsigned m = 0;
  for (int k = threadIdx.x; k < 4096; k += 1024) { m = max(m, __ldcg(g_cnt + k)); g_cnt[k] = 0u; }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) atomicMax(&s_max, m);
  unsigned ec = 0;
  if (threadIdx.x < 256) { ec = __ldcg(g_ecnt + threadIdx.x); g_ecnt[threadIdx.x] = 0u; if (ec) atomicMax(&s_emax, (unsigned)threadIdx.x); }
  __syncthreads();
  if (threadIdx.x < 256 && ec && s_emax > 14u && (unsigned)threadIdx.x < s_emax - 14u) atomicAdd(&s_low, ec);   // > 14 binades below the largest
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t_adj = __ldcg(g_misc);
    g_misc[0] = 0u; g_misc[1] = 0u;
    // 2: labels come in runs (sorted / contiguous) -> segmented warp reduce; 1: skewed -> hot-key cache; 0: neither
    status[4] = t_adj > (unsigned)samples / 2 ? 2 : ((s_max > (unsigned)thresh || t_adj > (unsigned)samples / 8) ? 1 : 0);
    status[5] = (int)s_max;
    status[6] = (int)t_adj;
    // 1: more than 0.1 % of the sampled values lie below the two-word fixed-point window (16 binades, minus a margin)
    // of the largest one: the hot-key cache then gives every slot an f64 side accumulator
    status[7] = (a && s_low * 1024u > (unsigned)samples) ? 1 : 0;
  }
}

// kernel-selection knobs (agc_tuning(); the environment gives the initial values)
//   key 0  AGC_LARGE      strategy for tables larger than one SM's shared memory: 0 hot-key cache always, 1 the label
//                         probe decides (default), 3 always partial k_fast, 4 always the run kernel
//   key 1  AGC_CACHE_CTAS CTAs per SM of k_cache (1 | 2)
//   key 2  AGC_SUM1       large tables: one-word sums with carry forwarding / split mean (1, default) or the fused
//                         two- and three-word tables (0)
//   key 3  AGC_SMEM1      bytes of dynamic shared memory one CTA per SM may use for its table.  Default 199552: with
//                         more than 196 KB of shared memory per SM the L1 keeps 28 KB, which caps the global loads
//                         in flight -- a streaming kernel then reaches 350-370 instead of 500-590 Gelem/s
//                         (measured cliff between 196 000 and 200 000 B of table, profiles/r01d_smem_carveout_cliff.jsonl)
//   key 4  AGC_SMEM2      the same for kernels that run two CTAs per SM (default 99200)
//   key 5  AGC_UNCOV_PCT  a partial table keeps the 195 KB size while at most this percentage of the groups stays uncovered,
//                         beyond that it grows to 226 KB (default 10; measured crossover 9-17 %, scripts/uncov_sweep.py)
//   key 6  (agc_scan.cuh) sorted-label scans: chained single pass (1, default) or the three-kernel variant (0)
//   key 7  AGC_TEAM       sums / means on tables larger than one SM: thread-block clusters of this many SMs own the table
//                         together and exchange records over DSMEM (k_team, agc_team.cuh); 0 = partial table + global RED.
//                         Default 0: measured on B200, a record that travels through the exchange costs ~9 issue clocks per
//                         SM (compaction, st.async, receive loop) where a global RED costs ~2.6 (profiles/r02_team_*.jsonl)
//   key 8  AGC_TEAM_MAXU  k_team runs while at most this percentage of the groups stays without a slot (default 60)
//   key 9  AGC_PART_GRID  CTAs of the partial-table k_fast (0 = one per SM, default).  The receiving side of the scattered REDs
//                         peaks with 111-129 senders (ubench/tma_red_probe.cu: 151 G RED/s against 141 from 148 SMs)
enum { TUNE_LARGE = 0, TUNE_CACHE_CTAS = 1, TUNE_SUM1 = 2, TUNE_SMEM1 = 3, TUNE_SMEM2 = 4, TUNE_UNCOV = 5, TUNE_SCAN_CHAIN = 6, TUNE_TEAM = 7,
       TUNE_TEAM_MAXU = 8, TUNE_PART_GRID = 9, TUNE_COUNT = 10 };
static int g_tune[TUNE_COUNT] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
inline void tuning_defaults_once() {
  const char* e;
  e = getenv("AGC_LARGE");      g_tune[TUNE_LARGE] = e ? atoi(e) : 1;
  if (g_tune[TUNE_LARGE] != 0 && g_tune[TUNE_LARGE] != 3 && g_tune[TUNE_LARGE] != 4) g_tune[TUNE_LARGE] = 1;
  e = getenv("AGC_CACHE_CTAS"); g_tune[TUNE_CACHE_CTAS] = (e && atoi(e) == 1) ? 1 : 2;
  e = getenv("AGC_SUM1");       g_tune[TUNE_SUM1] = (e && atoi(e) == 0) ? 0 : 1;
  e = getenv("AGC_SMEM1");      g_tune[TUNE_SMEM1] = e ? atoi(e) : 199552;
  e = getenv("AGC_SMEM2");      g_tune[TUNE_SMEM2] = e ? atoi(e) : 99200;
  e = getenv("AGC_UNCOV_PCT");  g_tune[TUNE_UNCOV] = e ? atoi(e) : 10;
  g_tune[TUNE_SCAN_CHAIN] = 1;                                 // lives in agc_scan.cuh (agc_tuning forwards key 6)
  e = getenv("AGC_TEAM");       g_tune[TUNE_TEAM] = e ? atoi(e) : 0;
  if (g_tune[TUNE_TEAM] != 0 && g_tune[TUNE_TEAM] != 2 && g_tune[TUNE_TEAM] != 4) g_tune[TUNE_TEAM] = 0;
  e = getenv("AGC_TEAM_MAXU");  g_tune[TUNE_TEAM_MAXU] = e ? atoi(e) : 60;
  e = getenv("AGC_PART_GRID");  g_tune[TUNE_PART_GRID] = e ? atoi(e) : 0;
}
inline void tuning_defaults() { static std::once_flag once; std::call_once(once, tuning_defaults_once); }

// k_team launcher over its template grid; returns the number of clusters (> 0), or < 0 if the cluster launch is not possible
template <class LT>
inline int launch_team_lt(int C, int mode, bool nanskip, const TeamParams& tp, size_t smem, cudaStream_t st) {
#define AGC_TEAM_ONE(C_, M_)                                                                                          \
  if (C == C_ && mode == M_) return nanskip ? launch_team<C_, LT, M_, true>(tp, smem, st) : launch_team<C_, LT, M_, false>(tp, smem, st);
  AGC_TEAM_ONE(2, TM_SUM) AGC_TEAM_ONE(2, TM_SUMCNT) AGC_TEAM_ONE(4, TM_SUM) AGC_TEAM_ONE(4, TM_SUMCNT)
#undef AGC_TEAM_ONE
  return -33;
}
inline size_t smem_cap1() { int v = g_tune[TUNE_SMEM1]; return (size_t)(v < 16384 ? 16384 : (v > 227 * 1024 - 1024 ? 227 * 1024 - 1024 : v)); }
inline size_t smem_cap2() { int v = g_tune[TUNE_SMEM2]; return (size_t)(v < 8192 ? 8192 : (v > 110 * 1024 ? 110 * 1024 : v)); }

// second pass of var / std (T0 holds the means, T2 takes sum (a - mean)^2): shared-memory tables for f32 data and f32
// results while one SM holds a sum word per group.  Returns 1 if it ran, 0 to fall back to the general kernel.
inline int try_fast_pass2(const agc_request_t* r, const Tables& t, cudaStream_t st, int sms) {
  if (r->op != AGC_OP_VAR || r->index.form != AGC_FORM_FLAT) return 0;
  if (r->index.dtype != AGC_DT_I64 && r->index.dtype != AGC_DT_I32) return 0;
  if (r->a == nullptr || r->a_dtype != AGC_DT_F32 || r->out_dtype != AGC_DT_F32) return 0;
  if (r->n < (1 << 16) || r->size >= (1ll << 31)) return 0;
  if (((uintptr_t)r->index.labels & 15) || ((uintptr_t)r->a & 15)) return 0;
  tuning_defaults();
  if (g_tune[TUNE_SUM1] == 0) return 0;
  const size_t cap2 = smem_cap2(), cap1 = smem_cap1();
  const bool two_word = (size_t)r->size * 8 <= cap1;          // exact 16-binade window
  if (!two_word && (size_t)r->size * 4 > cap1) return 0;      // a partial table would need the label probe: general kernel
  FastParams p;
  p.labels = r->index.labels; p.a = (const float*)r->a; p.n = r->n; p.size = r->size; p.t = t; p.status = r->status;
  p.t.t0f = t.t2; p.means = t.t0f;
  p.flags = r->flags & AGC_F_NANSKIP; p.square = 0; p.want_b1 = 0; p.want_cnt = 0;
  p.rshift = 0; p.nslots = 0; p.choose = nullptr; p.want = 0; p.cov = (int)r->size; p.choose2 = nullptr; p.want2 = 0;
  const bool i64 = r->index.dtype == AGC_DT_I64;
  const size_t table = (size_t)r->size * (two_word ? 8 : 4);
  // up to 12.4 K groups the means ride along in shared memory (8 more bytes per group): an LDS instead of an L2 round
  // trip per element
  const size_t mbytes = (size_t)r->size * 8 + 8;
  p.means_smem = table + mbytes <= cap1 ? 1 : 0;
  const size_t extra = p.means_smem ? mbytes : 0;
  p.side = (two_word && 2 * table + extra <= cap1) ? 1 : 0;   // f64 side accumulator for what falls below the rounding window
  const size_t tbytes = p.side ? 2 * table : table;
  const int ctas = tbytes + extra <= cap2 ? 2 : 1;
  if (ctas == 2) while (p.rshift < 5 && (tbytes << (p.rshift + 1)) + extra <= cap2) ++p.rshift;
  const size_t smem = (tbytes << p.rshift) + extra;
  const long long per_launch = (long long)sms * ctas * FAST_MAX_PER_CTA;
  int rc = 1;
  for (long long off = 0; off < r->n && rc >= 0; off += per_launch) {
    p.n = r->n - off < per_launch ? r->n - off : per_launch;
    p.labels = (const char*)r->index.labels + off * (i64 ? 8 : 4);
    p.a = (const float*)r->a + off;
    if (two_word) rc = i64 ? launch_fast<FM_SUM, long long, true>(p, ctas, smem, st, sms) : launch_fast<FM_SUM, int, true>(p, ctas, smem, st, sms);
    else          rc = i64 ? launch_fast<FM_SUM1, long long, true>(p, ctas, smem, st, sms) : launch_fast<FM_SUM1, int, true>(p, ctas, smem, st, sms);
  }
  return rc < 0 ? rc : 1;
}

// returns 1 if it accumulated into the global tables (which it also initialises), 0 to fall back
inline int try_fast_path(const agc_request_t* r, const Tables& t, cudaStream_t st, int sms, unsigned* probe_scratch) {
  if (r->index.form == AGC_FORM_AXIS) return try_form3_path(r, t, st, sms);
  if (r->index.form != AGC_FORM_FLAT) return 0;
  if (r->index.dtype != AGC_DT_I64 && r->index.dtype != AGC_DT_I32) return 0;
  const int op = r->op;
  const bool is_len = op == AGC_OP_LEN;
  const bool nanskip = r->flags & AGC_F_NANSKIP;
  if (!(op == AGC_OP_SUM || op == AGC_OP_SUMSQ || op == AGC_OP_MEAN || op == AGC_OP_MIN || op == AGC_OP_MAX || is_len || op == AGC_OP_VAR ||
        op == AGC_OP_MINMAX))
    return 0;
  const bool need_a = !is_len || nanskip;
  if (need_a && (r->a == nullptr || r->a_dtype != AGC_DT_F32)) return 0;
  if (r->n < (1 << 16) || r->size >= (1ll << 31)) return 0;
  if (((uintptr_t)r->index.labels & 15) || (need_a && ((uintptr_t)r->a & 15))) return 0;
  int mode;
  const bool want_b1 = (op == AGC_OP_SUM || op == AGC_OP_SUMSQ) && (r->flags & AGC_F_NEED_TOUCHED);
  if (is_len) mode = FM_LEN;
  else if (op == AGC_OP_MAX) mode = FM_MAX;
  else if (op == AGC_OP_MIN) mode = FM_MIN;
  else if (op == AGC_OP_MINMAX) mode = FM_MINMAX;
  else if (op == AGC_OP_MEAN || op == AGC_OP_VAR || want_b1) mode = FM_SUMCNT;
  else mode = FM_SUM;
  const int words = (mode == FM_SUM || mode == FM_MINMAX) ? 2 : (mode == FM_SUMCNT ? 3 : 1);
  tuning_defaults();
  // fused min + max: two key words per group while the whole table fits one SM; larger tables run on the general kernel
  if (mode == FM_MINMAX && (size_t)r->size * 8 > smem_cap1()) return 0;
  const size_t cap2 = smem_cap2(), cap1 = smem_cap1();
  const size_t table = (size_t)r->size * words * 4;
  const bool direct = table <= cap1;          // whole table privatised per CTA
  const int env_cache_ctas = g_tune[TUNE_CACHE_CTAS];
  const int strategy = g_tune[TUNE_LARGE];
  const int uniform_kernel = (strategy == 1 || strategy == 3) ? 3 : (strategy == 4 ? -1 : 0);   // 3: partial-table k_fast
  const bool probe = strategy == 1;

  FastParams p;
  p.labels = r->index.labels; p.a = (const float*)r->a; p.n = r->n; p.size = r->size; p.t = t; p.status = r->status;
  p.flags = r->flags; p.square = op == AGC_OP_SUMSQ; p.want_b1 = want_b1;
  p.want_cnt = op == AGC_OP_MEAN || op == AGC_OP_VAR;
  p.rshift = 0; p.nslots = 0; p.choose = nullptr; p.want = 0; p.cov = (int)r->size; p.means = nullptr; p.means_smem = 0; p.side = 0; p.choose2 = nullptr; p.want2 = 0;
  AGC_COUNT_LAUNCH(1), k_init_tables<<<(int)((r->size + 255) / 256), 256, 0, st>>>(t, r->size, r->op, r->a_dtype, r->out_dtype, 0);
  const bool i64 = r->index.dtype == AGC_DT_I64;
  int rc = 0;

#define AGC_FAST_LAUNCH(LAUNCHER, M, ...)                                                                        \
  if (mode == M) {                                                                                               \
    if (i64) rc = need_a ? LAUNCHER<M, long long, true>(__VA_ARGS__) : LAUNCHER<M, long long, false>(__VA_ARGS__); \
    else     rc = need_a ? LAUNCHER<M, int, true>(__VA_ARGS__) : LAUNCHER<M, int, false>(__VA_ARGS__);             \
  }
#define AGC_FAST_ALL(LAUNCHER, ...)                                                                              \
  AGC_FAST_LAUNCH(LAUNCHER, FM_SUM, __VA_ARGS__) AGC_FAST_LAUNCH(LAUNCHER, FM_SUMCNT, __VA_ARGS__)                 \
  AGC_FAST_LAUNCH(LAUNCHER, FM_MAX, __VA_ARGS__) AGC_FAST_LAUNCH(LAUNCHER, FM_MIN, __VA_ARGS__) AGC_FAST_LAUNCH(LAUNCHER, FM_LEN, __VA_ARGS__)
#define AGC_FAST_DIRECT(LAUNCHER, ...) AGC_FAST_ALL(LAUNCHER, __VA_ARGS__) AGC_FAST_LAUNCH(LAUNCHER, FM_MINMAX, __VA_ARGS__)

  // a CTA may take at most FAST_MAX_PER_CTA elements per launch (fixed-point / u32-count headroom): chunk N
  auto chunked = [&](int ctas, size_t smem, bool cache) {
    const long long per_launch = (long long)sms * ctas * FAST_MAX_PER_CTA;
    const long long n_total = r->n;
    for (long long off = 0; off < n_total && rc >= 0; off += per_launch) {
      p.n = n_total - off < per_launch ? n_total - off : per_launch;
      p.labels = (const char*)r->index.labels + off * (i64 ? 8 : 4);
      p.a = need_a ? (const float*)r->a + off : nullptr;
      if (cache) { AGC_FAST_ALL(launch_cache, p, ctas, smem, st, sms) }
      else { AGC_FAST_DIRECT(launch_fast, p, ctas, smem, st, sms) }
    }
    p.n = r->n; p.labels = r->index.labels; p.a = (const float*)r->a;
  };

  // mean / var (and sums that report touched groups) on 16.6 K - 24.9 K groups: one-word sums + u32 counts, 8 B per group,
  // still fit one SM in one pass
  if (!direct && mode == FM_SUMCNT && g_tune[TUNE_SUM1] != 0 && (size_t)r->size * 8 <= cap1) {
    const long long per_launch = (long long)sms * FAST_MAX_PER_CTA;
    const size_t smem = (size_t)r->size * 8;
    for (long long off = 0; off < r->n && rc >= 0; off += per_launch) {
      p.n = r->n - off < per_launch ? r->n - off : per_launch;
      p.labels = (const char*)r->index.labels + off * (i64 ? 8 : 4);
      p.a = (const float*)r->a + off;
      if (i64) rc = launch_fast<FM_SUMCNT1, long long, true>(p, 1, smem, st, sms);
      else     rc = launch_fast<FM_SUMCNT1, int, true>(p, 1, smem, st, sms);
    }
    return rc;
  }

  if (direct) {
    // sums on small tables: 8 more bytes per entry for the f64 side accumulator (values outside the fixed-point window
    // stay on chip instead of queueing as REDs on a few hot addresses)
    size_t tbytes = table;
    if ((mode == FM_SUM || mode == FM_SUMCNT) && (size_t)r->size * (words * 4 + 8) <= cap1) { p.side = 1; tbytes = (size_t)r->size * (words * 4 + 8); }
    const int ctas = tbytes <= cap2 ? 2 : 1;
    if (ctas == 2) while (p.rshift < 5 && (tbytes << (p.rshift + 1)) <= cap2) ++p.rshift;   // replicate small tables
    chunked(ctas, (tbytes << p.rshift) + (p.side ? 8 : 0), false);        // + alignment pad in front of the side table
    return rc;
  }

  // ---- max / min up to 113 K groups: the 16-bit filter kernel for uniform and skewed labels alike; labels in runs
  //      (the probe says 2) go to the segmented warp reduce, which streams them at HBM rate ------------------------
  // Partial tables: how many groups get a shared-memory slot of `bpg` bytes.  Up to 195 KB the SM keeps 60 KB of L1 and the
  // kernel streams at HBM rate; the last 31 KB of shared memory cost a third of the load bandwidth, which only pays when
  // the kernel is bound by the REDs of its uncovered groups anyway (measured: profiles/r01d_smem_carveout_cliff.jsonl).
  const size_t cap_max = 227 * 1024 - 1024;
  auto pick_cov = [&](size_t bpg_num, size_t bpg_den) -> long long {      // bytes per group = bpg_num / bpg_den
    const long long fast = (long long)(cap1 * bpg_den / bpg_num), big = (long long)(cap_max * bpg_den / bpg_num);
    if (r->size <= fast) return r->size;
    if (cap1 >= cap_max || (double)(r->size - fast) <= 0.01 * g_tune[TUNE_UNCOV] * (double)r->size) return fast & ~1ll;
    return (r->size < big ? r->size : big) & ~1ll;
  };
  const long long cov16 = pick_cov(2, 1);                    // groups a table of 16-bit entries covers
  if ((mode == FM_MAX || mode == FM_MIN) && g_tune[TUNE_LARGE] == 1 && g_tune[TUNE_SUM1] != 0 && r->size <= 8 * (long long)(cap_max / 2)) {
    long long samples = r->n / 256;
    samples = samples < 4096 ? 4096 : (samples > PROBE_SAMPLES ? PROBE_SAMPLES : samples);
    samples &= ~1023ll;
    const int thresh = (int)(12 + samples * (PROBE_SKEW_COUNT - 12) / PROBE_SAMPLES);
    cudaMemsetAsync(probe_scratch, 0, (size_t)PROBE_SCRATCH_WORDS * 4, st);       // the workspace is the caller's: nothing in it is zero yet
    if (i64) k_skew_probe<long long><<<(int)(samples / 1024), 1024, 0, st>>>(r->index.labels, r->n, r->status, (int)samples, thresh, probe_scratch);
    else k_skew_probe<int><<<(int)(samples / 1024), 1024, 0, st>>>(r->index.labels, r->n, r->status, (int)samples, thresh, probe_scratch);
    AGC_COUNT_LAUNCH(1);
    p.choose = r->status + 4; p.want = 0;
    // labels the probe calls uniform: groups past the filter table always pass (one RED each; no group holds more than
    // ~0.1 % of the elements, so none of them can pile more than ~1.5 ms of REDs on one address)
    p.cov = (int)(r->size < cov16 ? r->size : cov16);
    if (mode == FM_MAX) rc = i64 ? launch_maxfilter<true, long long, false>(p, st, sms) : launch_maxfilter<true, int, false>(p, st, sms);
    else rc = i64 ? launch_maxfilter<false, long long, false>(p, st, sms) : launch_maxfilter<false, int, false>(p, st, sms);
    if (rc < 0) return rc;
    // skewed labels: a hot group must never be one of the uncovered ones (ONE such group with 1.5 % of the elements costs
    // 18 ms of serialised REDs).  Up to 113 K groups the whole table fits (in the 228 KB carve-out: slower streaming, no
    // hazard) and ties get the exact test; beyond that the hot-key cache takes over.
    p.want = 1;
    if ((size_t)((r->size + 1) / 2) * 4 <= cap_max) {
      p.cov = (int)r->size;
      if (mode == FM_MAX) rc = i64 ? launch_maxfilter<true, long long, true>(p, st, sms) : launch_maxfilter<true, int, true>(p, st, sms);
      else rc = i64 ? launch_maxfilter<false, long long, true>(p, st, sms) : launch_maxfilter<false, int, true>(p, st, sms);
    } else {
      const int ctas = env_cache_ctas;
      const size_t budget = ctas == 2 ? 110 * 1024 : cap_max;
      p.cov = (int)r->size;
      p.nslots = (int)(budget / ((1 + words) * 4)) & ~1;
      chunked(ctas, (size_t)p.nslots * (1 + words) * 4, true);
    }
    if (rc < 0) return rc;
    p.want = 2; p.cov = (int)r->size;
    AGC_FAST_ALL(launch_runs, p, st, sms)
    return rc;
  }

  // ---- table larger than one SM's shared memory ---------------------------------------------------------
  // non-skewed labels: partial table + global RED (k_fast with cov < size);
  // skewed labels / runs: per-CTA hot-key cache (k_cache).  The on-device label probe picks; both kernels are
  // launched and the one that was not chosen exits at once (no host synchronisation).
  int* choose = r->status + 4;
  if (probe) {
    // the probe is one CTA and costs ~1 us per 1024 samples: scale it with the input (a hot group hurts in proportion to n)
    long long samples = r->n / 256;
    samples = samples < 4096 ? 4096 : (samples > PROBE_SAMPLES ? PROBE_SAMPLES : samples);
    samples &= ~1023ll;
    const int thresh = (int)(12 + samples * (PROBE_SKEW_COUNT - 12) / PROBE_SAMPLES);       // 40 at 32 Ki samples (0.1 % share), 15 at 4 Ki (0.35 %)
    const float* pa = need_a ? (const float*)r->a : nullptr;
    cudaMemsetAsync(probe_scratch, 0, (size_t)PROBE_SCRATCH_WORDS * 4, st);
    if (i64) k_skew_probe<long long><<<(int)(samples / 1024), 1024, 0, st>>>(r->index.labels, r->n, r->status, (int)samples, thresh, probe_scratch, pa);
    else k_skew_probe<int><<<(int)(samples / 1024), 1024, 0, st>>>(r->index.labels, r->n, r->status, (int)samples, thresh, probe_scratch, pa);
    AGC_COUNT_LAUNCH(1);
  }
  if (uniform_kernel == 3) {
    // (partial) table in shared memory: groups [0, cov), the rest global RED.  Plain sums use the one-word format
    // FM_SUM1 here (twice the coverage; its carries are REDs too, which is why it needs the probe's "not skewed").
    // (Measured and rejected: CTA pairs that stream the same tiles and split the covered groups -- max at 10^5
    // groups 227 vs 285 Gelem/s.)
    const bool sum1 = mode == FM_SUM && g_tune[TUNE_SUM1] != 0;
    // mean / var pass 1: one-word sums + a separate count pass over the labels (20 B/element, two kernels) beat the fused
    // three-word table, whose coverage is a third and whose every uncovered element costs TWO RED packets
    const bool split = mode == FM_SUMCNT && p.want_cnt && !want_b1 && g_tune[TUNE_SUM1] != 0 &&
                       r->size * 4 > (long long)(cap1 / 12) * 5;   // the fused table would cover < 80 % of the groups
    p.choose = probe ? choose : nullptr; p.want = 0;
    // one launch group of k_fast<M> with `w` words per group; `with_a`: the kernel reads the values
    const int sms_all = sms;
    auto run_fast = [&](int m, int w, bool with_a, long long max_per_cta) {
      p.cov = (int)pick_cov((size_t)w * 4, 1);
      const int sms = (p.cov < p.size && g_tune[TUNE_PART_GRID] > 0 && g_tune[TUNE_PART_GRID] < sms_all) ? g_tune[TUNE_PART_GRID] : sms_all;
      const long long per_launch = (long long)sms * max_per_cta;
      for (long long off = 0; off < r->n && rc >= 0; off += per_launch) {
        p.n = r->n - off < per_launch ? r->n - off : per_launch;
        p.labels = (const char*)r->index.labels + off * (i64 ? 8 : 4);
        p.a = with_a ? (const float*)r->a + off : nullptr;
        const size_t smem = (size_t)p.cov * w * 4;
#define AGC_RUN_FAST(M)                                                                                          \
        if (m == M) {                                                                                            \
          if (i64) rc = with_a ? launch_fast<M, long long, true>(p, 1, smem, st, sms) : launch_fast<M, long long, false>(p, 1, smem, st, sms); \
          else     rc = with_a ? launch_fast<M, int, true>(p, 1, smem, st, sms) : launch_fast<M, int, false>(p, 1, smem, st, sms);             \
        }
        AGC_RUN_FAST(FM_SUM) AGC_RUN_FAST(FM_SUMCNT) AGC_RUN_FAST(FM_MAX) AGC_RUN_FAST(FM_MIN) AGC_RUN_FAST(FM_LEN) AGC_RUN_FAST(FM_SUM1)
#undef AGC_RUN_FAST
      }
      p.n = r->n; p.labels = r->index.labels; p.a = (const float*)r->a;
    };
    const long long no_limit = FAST_MAX_PER_CTA * 64;           // FM_SUM1 forwards its carries: no fixed-point headroom to respect
    // counts: 16-bit packed counters hold the whole table up to 113 K groups (k_len16), else the u32 partial table
    const bool len16 = g_tune[TUNE_SUM1] != 0 && r->size <= 8 * (long long)(cap_max / 2);
    auto run_len = [&]() {
      if (len16) {
        p.cov = (int)(r->size < cov16 ? r->size : cov16);
        if (i64) rc = nanskip ? launch_len16<long long, true>(p, st, sms) : launch_len16<long long, false>(p, st, sms);
        else     rc = nanskip ? launch_len16<int, true>(p, st, sms) : launch_len16<int, false>(p, st, sms);
      } else run_fast(FM_LEN, 1, nanskip, no_limit);             // u32 counts: < 2^32 elements per CTA
    };
    // ---- k_team: a cluster of SMs owns the table together (no scattered RED for the groups it covers) -------------
    bool team_done = false;
    const int team_c = g_tune[TUNE_TEAM];
    if (team_c && need_a && !p.square && (mode == FM_SUM || mode == FM_SUMCNT) && g_tune[TUNE_SUM1] != 0) {
      const int tmode = mode == FM_SUM ? TM_SUM : TM_SUMCNT;
      const int tcap = team_default_cap(team_c);
      const long long fixed = (long long)team_smem_bytes(team_c, tmode, 0, tcap) + 16;
      const long long slots = ((long long)cap1 - fixed) / (tmode == TM_SUM ? 4 : 6);
      long long cov = r->size / team_c;                         // cov * C <= size (the kernel relies on it)
      if (cov > slots) cov = slots;
      const double uncovered = 1.0 - (double)(cov * team_c) / (double)r->size;
      if (cov >= 1024 && uncovered * 100.0 <= (double)g_tune[TUNE_TEAM_MAXU]) {
        TeamParams tp;
        tp.labels = r->index.labels; tp.a = (const float*)r->a; tp.n = r->n; tp.size = r->size; tp.t = t; tp.status = r->status;
        tp.flags = r->flags; tp.want_cnt = p.want_cnt; tp.want_b1 = p.want_b1; tp.cov = (int)cov; tp.cap = tcap;
        tp.choose = probe ? choose : nullptr; tp.want = 0;
        const size_t tsmem = team_smem_bytes(team_c, tmode, (int)cov, tcap);
        const int trc = i64 ? launch_team_lt<long long>(team_c, tmode, nanskip, tp, tsmem, st) : launch_team_lt<int>(team_c, tmode, nanskip, tp, tsmem, st);
        if (trc > 0) team_done = true;
        else cudaGetLastError();                                // clusters not available: the round-1 kernels below take over
      }
    }
    if (team_done) {}
    else if (sum1) run_fast(FM_SUM1, 1, true, no_limit);
    else if (split) {
      run_fast(FM_SUM1, 1, true, no_limit);
      if (rc >= 0) run_len();
    } else if (mode == FM_LEN) run_len();
    else run_fast(mode, words, need_a, FAST_MAX_PER_CTA);
    if (rc < 0) return rc;
    p.cov = (int)r->size;
  }
  if (strategy == 4 || probe) {                             // labels in runs: segmented warp reduce, no table at all
    p.choose = probe ? choose : nullptr; p.want = 2; p.cov = (int)r->size;
    AGC_FAST_ALL(launch_runs, p, st, sms)
    if (rc < 0) return rc;
  }
  if (uniform_kernel == 0 || probe) {
    const int ctas = env_cache_ctas;
    const size_t budget = ctas == 2 ? 110 * 1024 : cap_max;     // the cache is RED-bound on its misses: capacity beats L1 here
    const bool is_sum = mode == FM_SUM || mode == FM_SUMCNT;
    p.choose = probe ? choose : nullptr; p.want = 1;
    // sums: with the side accumulators (8 more bytes per slot) when the probe saw values spread over more binades than the
    // fixed-point window holds, without them (more slots, higher hit rate) for narrow data; no probe: with
    for (int with_side = (is_sum && probe) ? 0 : (is_sum ? 1 : 0); with_side <= (is_sum ? 1 : 0) && rc >= 0; ++with_side) {
      const size_t slot_b = (1 + words) * 4 + (with_side ? 8 : 0);
      p.side = with_side;
      p.nslots = (int)(budget / slot_b) & ~1;                   // even: slot s ^ 1 is the second way
      if (is_sum && probe) { p.choose2 = r->status + 7; p.want2 = with_side; }
      chunked(ctas, (size_t)p.nslots * slot_b, true);
    }
    p.choose2 = nullptr; p.side = 0;
  }
#undef AGC_FAST_DIRECT
#undef AGC_FAST_ALL
#undef AGC_FAST_LAUNCH
  return rc < 0 ? rc : 1;
}

}  // namespace agc
