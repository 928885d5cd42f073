// agc_labels.cuh -- the label utilities that sit either side of `aggregate` in real pipelines:
//   step_count / step_indices   aggregate_numba.py:577-605   run boundaries of group_idx
//   label_contiguous_1d         utils.py:597-632             label contiguous blocks of a mask / of equal values
//   relabel_groups_masked       utils.py:651-681             compress group numbers after removing groups
//   relabel_groups_unique       utils.py:635-648             ... after removing the groups that never occur
//   multi_arange                utils.py:572-594             hstack(arange(n_i) for n_i in n)
// All of them are "one flag / one integer per element -> running sum -> map or compact", so they share ONE kernel:
// a single-pass chained scan (tile descriptors + warp-parallel look-back, agc_scan.cuh::chain_lookback) over int64.
#pragma once
#include "agc_scan.cuh"

namespace agc {

enum {
  LB_STEPS = 0,        // v_i = (i == 0 || x_i != x_{i-1});            out[excl_i] = i where v_i, out[total] = n   (step_indices)
  LB_STEP_COUNT = 1,   // same v_i, only the total                                                               (step_count)
  LB_CONTIG = 2,       // v_i = start of a block (bool: rising edge; ints: value change and non-zero);  out[i] = x_i ? incl_i : 0
  LB_KEEP = 3,         // v_k = keep_k (k = 0 always kept);            out[k] = keep_k ? excl_k : 0  in dtype out_dt    (relabel table)
  LB_CUMSUM = 4        // v_i = x_i;                                   out[i] = incl_i (int64), total                  (multi_arange)
};
constexpr int LB_THREADS = 256, LB_ITEMS = 8, LB_TILE = LB_THREADS * LB_ITEMS;

struct LabelScanParams {
  const void* x; int x_dtype;      // input (labels / mask / keep flags / counts)
  long long n;
  void* out; int out_dtype;        // LB_KEEP: any integer dtype; otherwise int64
  long long* total;                // device int64: sum of all v_i (may be NULL)
  unsigned long long* desc;        // ntiles x 16 B, zeroed
};

template <int MODE>
__global__ void __launch_bounds__(LB_THREADS) k_label_scan(const LabelScanParams p) {
  __shared__ long long s_warp[LB_THREADS / 32];
  __shared__ long long s_carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long base = (long long)blockIdx.x * LB_TILE + (long long)threadIdx.x * LB_ITEMS;
  long long x[LB_ITEMS], v[LB_ITEMS];
  long long prev = 0;
  if ((MODE == LB_STEPS || MODE == LB_STEP_COUNT || MODE == LB_CONTIG) && base > 0 && base - 1 < p.n) prev = load_int(p.x, p.x_dtype, base - 1);
  long long run = 0;
#pragma unroll
  for (int k = 0; k < LB_ITEMS; ++k) {
    const long long i = base + k;
    x[k] = 0; v[k] = 0;
    if (i < p.n) {
      x[k] = load_int(p.x, p.x_dtype, i);
      if (MODE == LB_STEPS || MODE == LB_STEP_COUNT) v[k] = (i == 0 || x[k] != prev);
      else if (MODE == LB_CONTIG) {
        // bool masks: a block starts on a rising edge; other dtypes: on every change of value, zeros never start one
        if (p.x_dtype == AGC_DT_BOOL) v[k] = (x[k] != 0) && (i == 0 || prev == 0);
        else v[k] = (x[k] != 0) && (i == 0 || x[k] != prev);
      } else if (MODE == LB_KEEP) v[k] = (i == 0 || x[k] != 0);
      else v[k] = x[k];
      prev = x[k];
    }
    run += v[k];
    v[k] = run;                                   // thread-local inclusive sums
  }
  // block scan of the thread totals
  long long incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  long long wbase = 0, tile_total = 0;
#pragma unroll
  for (int k = 0; k < LB_THREADS / 32; ++k) { const long long s = s_warp[k]; if (k < w) wbase += s; tile_total += s; }
  const long long excl_thread = wbase + incl - run;
  // chained carry
  if (threadIdx.x < 32) {
    const int t = blockIdx.x;
    if (lane == 0) desc_store(p.desc + 2 * (long long)t, t == 0 ? 2ull : 1ull, (unsigned long long)tile_total);
    long long c = 0;
    if (t > 0) {
      c = chain_lookback<long long, SC_SUM>(p.desc, t, lane);
      if (lane == 0) desc_store(p.desc + 2 * (long long)t, 2ull, (unsigned long long)(c + tile_total));
    }
    if (lane == 0) {
      s_carry = c;
      if (p.total && (long long)(blockIdx.x + 1) * LB_TILE >= p.n) *p.total = c + tile_total;    // last tile
    }
  }
  __syncthreads();
  const long long carry = s_carry + excl_thread;
  if (MODE == LB_STEP_COUNT) return;
#pragma unroll
  for (int k = 0; k < LB_ITEMS; ++k) {
    const long long i = base + k;
    if (i >= p.n) break;
    const long long inc = carry + v[k];
    const long long exc = carry + (k ? v[k - 1] : 0);
    if (MODE == LB_STEPS) { if (inc != exc) ((long long*)p.out)[exc] = i; }
    else if (MODE == LB_CONTIG) ((long long*)p.out)[i] = x[k] != 0 ? inc : 0;
    else if (MODE == LB_KEEP) store_from_i64(p.out, p.out_dtype, i, inc != exc ? exc : 0);
    else ((long long*)p.out)[i] = inc;
  }
  if (MODE == LB_STEPS && (long long)(blockIdx.x + 1) * LB_TILE >= p.n && threadIdx.x == 0) ((long long*)p.out)[s_carry + tile_total] = p.n;
}

inline int64_t label_scan_workspace_bytes(int64_t n) { return ((n + LB_TILE - 1) / LB_TILE + 1) * 16 + 256; }

inline int run_label_scan(int mode, const LabelScanParams& p0, void* workspace, cudaStream_t st) {
  LabelScanParams p = p0;
  if (p.n <= 0) { if (p.total) cudaMemsetAsync(p.total, 0, 8, st); return 0; }
  const long long ntiles = (p.n + LB_TILE - 1) / LB_TILE;
  if (ntiles > 0x7fffffffll) return -21;
  p.desc = (unsigned long long*)(((uintptr_t)workspace + 15) & ~(uintptr_t)15);
  cudaMemsetAsync(p.desc, 0, (size_t)ntiles * 16, st);
  AGC_COUNT_LAUNCH(1);
  switch (mode) {
    case LB_STEPS: k_label_scan<LB_STEPS><<<(int)ntiles, LB_THREADS, 0, st>>>(p); break;
    case LB_STEP_COUNT: k_label_scan<LB_STEP_COUNT><<<(int)ntiles, LB_THREADS, 0, st>>>(p); break;
    case LB_CONTIG: k_label_scan<LB_CONTIG><<<(int)ntiles, LB_THREADS, 0, st>>>(p); break;
    case LB_KEEP: k_label_scan<LB_KEEP><<<(int)ntiles, LB_THREADS, 0, st>>>(p); break;
    case LB_CUMSUM: k_label_scan<LB_CUMSUM><<<(int)ntiles, LB_THREADS, 0, st>>>(p); break;
    default: return -2;
  }
  return 0;
}

// keep[g] = 1 for every label that occurs (relabel_groups_unique, utils.py:645-647); bad labels raise status[0]
__global__ void k_mark_present(const void* labels, int dtype, long long n, long long size, unsigned char* keep, int* status) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  bool bad = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long g = load_int(labels, dtype, i);
    if (g < 0 || g >= size) { bad = true; continue; }
    if (!keep[g]) keep[g] = 1;
  }
  if (bad) status[AGC_ST_BADLABEL] = 1;
}

// multi_arange: out[j] = j - start of the segment that holds j; `incl` = inclusive running sums of n (m entries).
// Segment of j = first s with incl[s] > j (zero-length segments are skipped by the strict comparison).
__global__ void k_multi_arange(const long long* incl, long long m, long long total, long long* out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < total; j += stride) {
    long long lo = 0, hi = m - 1;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (incl[mid] > j) hi = mid; else lo = mid + 1; }
    out[j] = j - (lo ? incl[lo - 1] : 0);
  }
}

}  // namespace agc
