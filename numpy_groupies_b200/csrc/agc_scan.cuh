// agc_scan.cuh -- label utilities (min/max, unpack) and the N->N family:
//   * stable LSD radix sort-by-key (8-bit digits; keys = flat group or value key, payload = position)
//   * segmented inclusive scans over (sorted key, value) for cumsum / cumprod / cummax / cummin
//   * func='sort' (values ordered inside each group, written back to the group's own slots)
//   * group order + CSR offsets for func='array' and generic callables
// Reference semantics: aggregate_numpy.py:210-266 and aggregate_numba.py:218-225,487-500.
// Sortedness of the flat labels is detected on the device; already-sorted labels skip every sort
// pass without a host round trip (the sort kernels read the flag and return).
#pragma once
#include <cuda_fp16.h>
#include "agc_generic.cuh"

namespace agc {

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
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

// float16 <-> float32 (the kernels compute float16 data as float32: exact on the way in, round-to-nearest-even on the
// way out, the same as numpy's astype -- check_dtype hands float16 results back as float16, utils.py:435-470).
// 8 elements per thread-iteration: one 128-bit access on the half side, two on the float side.
__device__ __forceinline__ unsigned f16x2_bits(float lo, float hi) {
  return (unsigned)__half_as_ushort(__float2half_rn(lo)) | ((unsigned)__half_as_ushort(__float2half_rn(hi)) << 16);
}
__device__ __forceinline__ float f16_lo(unsigned w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
__device__ __forceinline__ float f16_hi(unsigned w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
template <bool TO_HALF>
__global__ void k_convert_f16(const void* __restrict__ src, void* __restrict__ dst, long long n) {
  const long long nvec = n >> 3, stride = (long long)gridDim.x * blockDim.x;
  const bool vec_ok = (((unsigned long long)src | (unsigned long long)dst) & 15ull) == 0;
  long long done = 0;
  if (vec_ok) {
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += stride) {
      if (TO_HALF) {
        const float4 a = ((const float4*)src)[2 * v], b = ((const float4*)src)[2 * v + 1];
        uint4 o;
        o.x = f16x2_bits(a.x, a.y); o.y = f16x2_bits(a.z, a.w); o.z = f16x2_bits(b.x, b.y); o.w = f16x2_bits(b.z, b.w);
        ((uint4*)dst)[v] = o;
      } else {
        const uint4 raw = ((const uint4*)src)[v];
        ((float4*)dst)[2 * v] = make_float4(f16_lo(raw.x), f16_hi(raw.x), f16_lo(raw.y), f16_hi(raw.y));
        ((float4*)dst)[2 * v + 1] = make_float4(f16_lo(raw.z), f16_hi(raw.z), f16_lo(raw.w), f16_hi(raw.w));
      }
    }
    done = nvec << 3;
  }
  for (long long i = done + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    if (TO_HALF) ((__half*)dst)[i] = __float2half_rn(((const float*)src)[i]);
    else ((float*)dst)[i] = __half2float(((const __half*)src)[i]);
  }
}

static thread_local const char* g_scan_err = "";
inline const char* scan_error() { return g_scan_err; }

// ---------------------------------------------------------------------------------------------
// device-wide exclusive scan of u32 counters (3 kernels, in place)
// ---------------------------------------------------------------------------------------------
constexpr int XS_THREADS = 256, XS_ITEMS = 16, XS_TILE = XS_THREADS * XS_ITEMS;

__device__ __forceinline__ unsigned block_exclusive_scan_u32(unsigned x, unsigned* total) {
  __shared__ unsigned wsum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned inc = x;
  for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned s = lane < nw ? wsum[lane] : 0, si = s;
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
    wsum[lane] = si - s;
    if (lane == 31) *total = si;
  }
  __syncthreads();
  unsigned r = wsum[w] + inc - x;
  __syncthreads();
  return r;
}
// (all tile kernels below walk their tiles with a grid-stride loop over a grid capped at a few CTAs per SM: when a
// device-side run_flag turns them into no-ops, an empty launch then costs microseconds instead of tens of them)
__global__ void k_xs_reduce(const unsigned* v, long long n, unsigned* tile_sums, int ntiles, const int* run_flag) {
  if (run_flag && *run_flag == 0) return;
  __shared__ unsigned tot;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long base = (long long)tile * XS_TILE + (long long)threadIdx.x * XS_ITEMS;
    unsigned s = 0;
    for (int k = 0; k < XS_ITEMS; ++k) if (base + k < n) s += v[base + k];
    block_exclusive_scan_u32(s, &tot);
    if (threadIdx.x == 0) tile_sums[tile] = tot;
    __syncthreads();
  }
}
__global__ void k_xs_scan_sums(unsigned* tile_sums, int ntiles, const int* run_flag) {      // single block, serial-in-thread chunks
  if (run_flag && *run_flag == 0) return;
  __shared__ unsigned tot;
  const int per = (ntiles + blockDim.x - 1) / blockDim.x;
  const int b = threadIdx.x * per;
  unsigned s = 0;
  for (int k = 0; k < per; ++k) if (b + k < ntiles) s += tile_sums[b + k];
  unsigned run = block_exclusive_scan_u32(s, &tot);
  for (int k = 0; k < per; ++k) if (b + k < ntiles) { unsigned t = tile_sums[b + k]; tile_sums[b + k] = run; run += t; }
}
__global__ void k_xs_apply(unsigned* v, long long n, const unsigned* tile_sums, int ntiles, const int* run_flag) {
  if (run_flag && *run_flag == 0) return;
  __shared__ unsigned tot;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long base = (long long)tile * XS_TILE + (long long)threadIdx.x * XS_ITEMS;
    unsigned x[XS_ITEMS], s = 0;
    for (int k = 0; k < XS_ITEMS; ++k) { x[k] = base + k < n ? v[base + k] : 0; s += x[k]; }
    unsigned run = block_exclusive_scan_u32(s, &tot) + tile_sums[tile];
    for (int k = 0; k < XS_ITEMS; ++k) if (base + k < n) { v[base + k] = run; run += x[k]; }
  }
}
inline long long xs_tiles(long long n) { return (n + XS_TILE - 1) / XS_TILE; }
inline void exclusive_scan_u32(unsigned* v, long long n, unsigned* tile_sums, cudaStream_t st, const int* run_flag = nullptr) {
  int nt = (int)xs_tiles(n);
  AGC_COUNT_LAUNCH(1), k_xs_reduce<<<tile_grid(nt, 8), XS_THREADS, 0, st>>>(v, n, tile_sums, nt, run_flag);
  AGC_COUNT_LAUNCH(1), k_xs_scan_sums<<<1, 1024, 0, st>>>(tile_sums, nt, run_flag);
  AGC_COUNT_LAUNCH(1), k_xs_apply<<<tile_grid(nt, 8), XS_THREADS, 0, st>>>(v, n, tile_sums, nt, run_flag);
}

// ---------------------------------------------------------------------------------------------
// radix sort-by-key: one 8-bit pass = tile histograms -> device-wide scan -> stable scatter.
// A tile is RS_THREADS x ITEMS consecutive elements; warp w owns the w-th contiguous slice of it and walks it in
// items of 32 consecutive elements, so (tile, warp, item, lane) order == element order and ranks are stable:
//   rank inside an item   MATCH.ANY on the digit (peers before me),
//   rank inside the warp  per-warp digit counters in shared memory (only the owning warp touches them: no block sync),
//   rank inside the tile  one pass over the 256 digits x warps after a single __syncthreads.
// The tile is then written to shared memory in digit order and copied out from there, so every store instruction
// covers whole runs of a digit (consecutive addresses) instead of 32 scattered words.
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 512, RS_WARPS = RS_THREADS / 32;
template <class K> struct RsCfg { static constexpr int ITEMS = 8; static constexpr int TILE = RS_THREADS * ITEMS; };   // 16 items -> 128 registers, 1 CTA/SM: slower
constexpr int RS_TILE_MIN = RS_THREADS * 8;                 // smallest tile of any key type: sizes the histogram workspace
inline int rs_blocks(long long n) { return (int)((n + RS_TILE_MIN - 1) / RS_TILE_MIN); }
template <class K> inline int rs_blocks_k(long long n) { return (int)((n + RsCfg<K>::TILE - 1) / RsCfg<K>::TILE); }

template <class K>
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const K* keys, long long n, int shift, unsigned* hist, int nblocks, const int* run_flag,
                                                           const K* bias_p = nullptr) {
  if (run_flag && *run_flag == 0) return;
  const K bias = bias_p ? *bias_p : K(0);                      // keys are sorted as (key - bias): see value_sort_range
  __shared__ unsigned h[256];
  for (int tile = blockIdx.x; tile < nblocks; tile += gridDim.x) {
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)tile * RsCfg<K>::TILE;
#pragma unroll 4
    for (int c = 0; c < RsCfg<K>::ITEMS; ++c) {
      const long long i = base + c * RS_THREADS + threadIdx.x;
      if (i < n) atomicAdd(&h[(unsigned)((K)(keys[i] - bias) >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256) hist[(long long)threadIdx.x * nblocks + tile] = h[threadIdx.x];
    __syncthreads();
  }
}

template <class K, class V = unsigned>
__global__ void __launch_bounds__(RS_THREADS, 2) k_radix_scatter(const K* kin, const V* vin, K* kout, V* vout, long long n, int shift,
                                                             const unsigned* offs, int nblocks, const int* run_flag, const K* bias_p = nullptr) {
  if (run_flag && *run_flag == 0) return;
  const K bias = bias_p ? *bias_p : K(0);
  constexpr int ITEMS = RsCfg<K>::ITEMS, TILE = RsCfg<K>::TILE;
  extern __shared__ __align__(16) unsigned char rs_smem[];
  V* sval = (V*)rs_smem;                                          // [TILE] tile in digit order (the wider type first: alignment)
  K* skey = (K*)(sval + TILE);                                    // [TILE]
  unsigned* cnt = (unsigned*)(skey + TILE);                       // [RS_WARPS][256] per-warp digit counters -> tile-local bases
  unsigned* dstart = cnt + RS_WARPS * 256;                        // [256] tile-local start of every digit
  unsigned* gbase = dstart + 256;                                 // [256] global position of the digit's first element of this tile
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int tile = blockIdx.x; tile < nblocks; tile += gridDim.x) {
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) cnt[i] = 0;
    __syncthreads();
    const long long tile0 = (long long)tile * TILE;
    const long long wbase = tile0 + (long long)w * (ITEMS * 32);
    K key[ITEMS]; V val[ITEMS]; unsigned short rnk[ITEMS];
    unsigned* mycnt = cnt + w * 256;
    const unsigned lt = (1u << lane) - 1u;
  #pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const long long i = wbase + j * 32 + lane;
      const bool valid = i < n;
      key[j] = valid ? kin[i] : K(0);
      val[j] = valid ? vin[i] : V(0);
    }
  #pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const bool valid = wbase + j * 32 + lane < n;
      const unsigned d = valid ? ((unsigned)((K)(key[j] - bias) >> shift) & 0xffu) : (256u + lane);
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const int leader = __ffs(peers) - 1;
      unsigned old = 0;
      if (lane == leader && valid) { old = mycnt[d]; mycnt[d] = old + __popc(peers); }
      old = __shfl_sync(0xffffffffu, old, leader);
      rnk[j] = (unsigned short)(old + __popc(peers & lt));
      __syncwarp();
    }
    __syncthreads();
    // digit d: exclusive scan over the warps, tile-local digit start, global base
    if (threadIdx.x < 256) {
      unsigned run = 0;
  #pragma unroll
      for (int k = 0; k < RS_WARPS; ++k) { const unsigned c = cnt[k * 256 + threadIdx.x]; cnt[k * 256 + threadIdx.x] = run; run += c; }
      dstart[threadIdx.x] = run;                                    // count for now
      gbase[threadIdx.x] = offs[(long long)threadIdx.x * nblocks + tile];
    }
    __syncthreads();
    if (w == 0) {                                                   // exclusive scan of the 256 digit counts (8 per lane)
      unsigned c[8], s = 0;
  #pragma unroll
      for (int k = 0; k < 8; ++k) { c[k] = dstart[lane * 8 + k]; s += c[k]; }
      unsigned inc = s;
  #pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      unsigned run = inc - s;
  #pragma unroll
      for (int k = 0; k < 8; ++k) { dstart[lane * 8 + k] = run; run += c[k]; }
    }
    __syncthreads();
  #pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (wbase + j * 32 + lane < n) {
        const unsigned d = (unsigned)((K)(key[j] - bias) >> shift) & 0xffu;
        const unsigned lp = dstart[d] + mycnt[d] + rnk[j];
        skey[lp] = key[j]; sval[lp] = val[j];
      }
    }
    __syncthreads();
    const int count = (int)(n - tile0 < TILE ? n - tile0 : TILE);
    for (int lp = threadIdx.x; lp < count; lp += RS_THREADS) {
      const K k = skey[lp];
      const unsigned d = (unsigned)((K)(k - bias) >> shift) & 0xffu;
      const long long pos = (long long)gbase[d] + (lp - dstart[d]);
      kout[pos] = k; vout[pos] = sval[lp];
    }
    __syncthreads();
  }
}
template <class K, class V = unsigned> inline size_t rs_scatter_smem() { return (size_t)RsCfg<K>::TILE * (sizeof(K) + sizeof(V)) + (RS_WARPS * 256 + 512) * 4; }

struct SortBufs {
  void* k[2]; unsigned* v[2]; unsigned* hist; unsigned* tile_sums;
};
// sorts bits [lo_bit, hi_bit) ; returns which buffer (0/1) holds the result. run_flag (device int, may be
// NULL) == 0 turns every kernel into a no-op, in which case the data stays in buffer `src`.
// pass_flags (device int per pass, may be NULL) replaces run_flag pass by pass; bias (device K, may be NULL) is subtracted
// from every key before its digit is taken.
template <class K>
inline int radix_sort_pairs(SortBufs& b, int src, long long n, int lo_bit, int hi_bit, const int* run_flag, cudaStream_t st,
                            const int* pass_flags = nullptr, const K* bias = nullptr) {
  if (n == 0) return src;
  const int nb = rs_blocks_k<K>(n);
  static std::atomic<unsigned long long> attr_done{0};
  ensure_dyn_smem(k_radix_scatter<K>, (int)rs_scatter_smem<K>(), attr_done);
  int pass = 0;
  for (int shift = lo_bit; shift < hi_bit; shift += 8, ++pass) {
    const int* rf = pass_flags ? pass_flags + pass : run_flag;
    AGC_COUNT_LAUNCH(1), k_radix_hist<K><<<tile_grid(nb, 4), RS_THREADS, 0, st>>>((const K*)b.k[src], n, shift, b.hist, nb, rf, bias);
    exclusive_scan_u32(b.hist, 256ll * nb, b.tile_sums, st, rf);
    AGC_COUNT_LAUNCH(1), k_radix_scatter<K><<<tile_grid(nb, 2), RS_THREADS, rs_scatter_smem<K>(), st>>>((const K*)b.k[src], b.v[src], (K*)b.k[src ^ 1], b.v[src ^ 1], n, shift,
                                                   b.hist, nb, rf, bias);
    src ^= 1;
  }
  return src;
}

// ---------------------------------------------------------------------------------------------
// key generation: flat group of every element (+ sortedness + bad-label detection)
// ---------------------------------------------------------------------------------------------
__global__ void k_group_keys(Indexer ix, long long n, unsigned* keys, unsigned* pos, int* status, int* unsorted, int only_if_unsorted) {
  if (only_if_unsorted && *unsorted == 0) return;          // the direct (sorted-label) scan already produced the result
  const long long stride = (long long)gridDim.x * blockDim.x;
  bool bad = false, uns = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    long long g = ix.group(i);
    if (g < 0) { bad = true; g = 0; }
    keys[i] = (unsigned)g; pos[i] = (unsigned)i;
    if (i > 0) { long long gp = ix.group(i - 1); if (gp > g) uns = true; }
  }
  if (bad) status[AGC_ST_BADLABEL] = 1;
  if (uns) *unsorted = 1;
}
// random gathers of single elements: ask L2 for the smallest DRAM fetch (64 bytes) instead of its default
__device__ __forceinline__ unsigned long long gather64(const void* base, unsigned long long idx) {
  unsigned long long v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"((const unsigned long long*)base + idx));
  return v;
}
__device__ __forceinline__ unsigned gather32(const void* base, unsigned long long idx) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b32 %0, [%1];" : "=r"(v) : "l"((const unsigned*)base + idx));
  return v;
}
__global__ void k_zero_int(int* p) { p[0] = 0; p[1] = 0; }
// 4096 neighbouring pairs spread over the labels: random labels show an inversion here with certainty, and the persistent
// direct scan (which finishes its pass once started) then never starts
template <class LT>
__global__ void k_sorted_sample(const LT* labels, long long n, int* unsorted) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x, m = (long long)gridDim.x * blockDim.x;
  if (n < 2) return;
  const long long i = (long long)((double)t / (double)m * (double)(n - 1));
  if (labels[i + 1] < labels[i]) { unsorted[0] = 1; unsorted[1] = 1; }
}
__global__ void k_report_unsorted(const int* unsorted, int* status) { if (*unsorted) status[AGC_ST_UNSORTED] = 1; }

// ---------------------------------------------------------------------------------------------
// segmented scans
// ---------------------------------------------------------------------------------------------
enum { SC_SUM = 0, SC_PROD = 1, SC_MAX = 2, SC_MIN = 3 };
constexpr int SG_THREADS = 256, SG_ITEMS = 8, SG_TILE = SG_THREADS * SG_ITEMS;

template <class A, int OP> struct ScanOp;
template <class A> struct ScanOp<A, SC_SUM> { static __device__ __forceinline__ A apply(A x, A y) { return x + y; } };
template <> struct ScanOp<long long, SC_SUM> {
  static __device__ __forceinline__ long long apply(long long x, long long y) { return (long long)((unsigned long long)x + (unsigned long long)y); } };
template <class A> struct ScanOp<A, SC_PROD> { static __device__ __forceinline__ A apply(A x, A y) { return x * y; } };
template <> struct ScanOp<long long, SC_PROD> {
  static __device__ __forceinline__ long long apply(long long x, long long y) { return (long long)((unsigned long long)x * (unsigned long long)y); } };
template <class A> struct ScanOp<A, SC_MAX> { static __device__ __forceinline__ A apply(A x, A y) { return x > y ? x : y; } };
template <class A> struct ScanOp<A, SC_MIN> { static __device__ __forceinline__ A apply(A x, A y) { return x < y ? x : y; } };

struct ScanParams {
  const unsigned* keys[2]; const unsigned* pos[2];    // [0] = as generated (sorted input), [1] = after the sort
  const int* unsorted;                                // selects between them on the device
  const void* a; int a_dtype; int out_dtype; void* out;
  long long n; int flags;
  double fill_f; long long fill_i;
  int a_float;                                        // value class of a
  int a_u64;
  const void* labels; long long size; int* status; int* unsorted_w;   // direct (sorted flat labels) path
  int only_if_unsorted;                               // keys path: no-op when the direct path already ran
  int lb_sleep;                                       // chained scan: nanoseconds a look-back lane sleeps before it polls an unpublished descriptor again
  int tile0;                                          // direct scan: tile of block 0 (the persistent chained kernel took the tiles before it)
  void* gath;                                         // keys path: scan inputs in sorted order (PHASE 0 gathers them once)
  void* stage;                                        // keys path, labels really unsorted: results in SORTED order (out dtype); the
                                                      // windowed un-sort below brings them home instead of one random store each
};

// element k of the sorted sequence -> the value fed to the scan (A = double for float sum/prod,
// long long otherwise; max/min always scan order-preserving integer keys)
template <class A, int OP>
__device__ __forceinline__ A scan_input(const ScanParams& p, unsigned src_pos, bool head) {
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  if (OP == SC_SUM || OP == SC_PROD) {
    if (std::is_same<A, double>::value) {
      double v;
      if (p.a_dtype == AGC_DT_F32) v = ((const float*)p.a)[src_pos];
      else if (p.a_dtype == AGC_DT_F64) v = __longlong_as_double((long long)gather64(p.a, src_pos));
      else if (p.a_u64) v = (double)gather64(p.a, src_pos);
      else v = (double)load_int(p.a, p.a_dtype, src_pos);
      if (nanskip && v != v) v = OP == SC_SUM ? 0.0 : 1.0;
      return (A)v;
    } else {
      if (p.a_dtype == AGC_DT_I64 || p.a_dtype == AGC_DT_U64) return (A)(long long)gather64(p.a, src_pos);
      return (A)load_int(p.a, p.a_dtype, src_pos);
    }
  } else {
    const long long ident = OP == SC_MAX ? AGC_EMPTY_MAX : AGC_EMPTY_MIN;
    const long long absorb = OP == SC_MAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX;
    if (p.a_float) {
      double v = p.a_dtype == AGC_DT_F32 ? (double)((const float*)p.a)[src_pos] : ((const double*)p.a)[src_pos];
      if (v != v) return (A)((nanskip || !head) ? ident : absorb);   // numba: a NaN only sticks when it opens the group
      return (A)key_of(v);
    }
    long long v = (p.a_dtype == AGC_DT_I64 || p.a_dtype == AGC_DT_U64) ? (long long)gather64(p.a, src_pos) : load_int(p.a, p.a_dtype, src_pos);
    return (A)(p.a_u64 ? key_of((unsigned long long)v) : v);
  }
}
template <class A, int OP>
__device__ __forceinline__ void scan_store(const ScanParams& p, long long dst, A r) {
  if (OP == SC_SUM || OP == SC_PROD) {
    if (std::is_same<A, double>::value) store_from_f64(p.out, p.out_dtype, dst, (double)r);
    else store_from_i64(p.out, p.out_dtype, dst, (long long)r, p.a_u64);
  } else {
    const long long ident = OP == SC_MAX ? AGC_EMPTY_MAX : AGC_EMPTY_MIN;
    const long long absorb = OP == SC_MAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX;
    long long k = (long long)r;
    if (p.a_float) {
      if (k == ident) { if (dtype_is_float(p.out_dtype)) store_from_f64(p.out, p.out_dtype, dst, p.fill_f); else store_from_i64(p.out, p.out_dtype, dst, p.fill_i); }
      else store_from_f64(p.out, p.out_dtype, dst, k == absorb ? __longlong_as_double(0x7ff8000000000000ll) : unkey_f64(k));
    } else {
      store_from_i64(p.out, p.out_dtype, dst, p.a_u64 ? (long long)((unsigned long long)k ^ 0x8000000000000000ull) : k, p.a_u64);
    }
  }
}

// block-wide segmented inclusive scan of per-thread aggregates (flag, value); returns the EXCLUSIVE
// prefix of this thread (has_prefix=false when nothing precedes it in the tile) and the tile aggregate.
template <class A, int OP>
__device__ __forceinline__ void block_seg_scan(bool f, A x, bool& pf, A& px, bool& have_p, bool& tile_f, A& tile_x) {
  __shared__ unsigned char s_f[SG_THREADS / 32];
  __shared__ long long s_x[SG_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  bool fi = f; A xi = x;                      // inclusive within warp
  for (int o = 1; o < 32; o <<= 1) {
    bool ft = __shfl_up_sync(0xffffffffu, (int)fi, o);
    A xt = __shfl_up_sync(0xffffffffu, xi, o);
    if (lane >= o) { if (!fi) xi = ScanOp<A, OP>::apply(xt, xi); fi = fi | ft; }
  }
  if (lane == 31) { s_f[w] = fi; ((A*)s_x)[w] = xi; }
  // exclusive within warp
  bool fe = __shfl_up_sync(0xffffffffu, (int)fi, 1); A xe = __shfl_up_sync(0xffffffffu, xi, 1);
  __syncthreads();
  // prefix over preceding warps (serial, <= 7 steps)
  bool wf = false; A wx = A(0); bool have_w = false;
  for (int k = 0; k < w; ++k) {
    bool kf = s_f[k]; A kx = ((A*)s_x)[k];
    if (!have_w) { wf = kf; wx = kx; have_w = true; }
    else { if (!kf) wx = ScanOp<A, OP>::apply(wx, kx); else wx = kx; wf = wf | kf; }
  }
  // combine: prefix = warps-before (+) lanes-before-in-warp
  if (lane == 0) { pf = wf; px = wx; have_p = have_w; }
  else {
    if (have_w && !fe) { px = ScanOp<A, OP>::apply(wx, xe); pf = wf | fe; }
    else { px = xe; pf = fe | wf; }
    have_p = true;
  }
  // tile aggregate = inclusive value of the last thread
  {
    bool tf = false; A tx = A(0); bool have = false;
    for (int k = 0; k < SG_THREADS / 32; ++k) {
      bool kf = s_f[k]; A kx = ((A*)s_x)[k];
      if (!have) { tf = kf; tx = kx; have = true; }
      else { if (!kf) tx = ScanOp<A, OP>::apply(tx, kx); else tx = kx; tf = tf | kf; }
    }
    tile_f = tf; tile_x = tx;
  }
  __syncthreads();
}

// PHASE 0: tile aggregates only.  PHASE 1: full scan with carry-in, writes the output.
template <class A, int OP, int PHASE>
__global__ void __launch_bounds__(SG_THREADS) k_seg_scan(const ScanParams p, unsigned char* tile_f, A* tile_x,
                                                         const unsigned char* carry_ok, const A* carry_x) {
  const int sel = *p.unsorted ? 1 : 0;
  if (p.only_if_unsorted && !sel) return;
  const unsigned* keys = p.keys[sel]; const unsigned* pos = p.pos[sel];
  const long long base = (long long)blockIdx.x * SG_TILE + (long long)threadIdx.x * SG_ITEMS;
  A x[SG_ITEMS]; bool h[SG_ITEMS]; unsigned ps[SG_ITEMS];
  unsigned prev = base > 0 && base - 1 < p.n ? keys[base - 1] : 0u;
  bool tf = false; A tx = A(0); bool any = false;
#pragma unroll
  for (int k = 0; k < SG_ITEMS; ++k) {
    long long i = base + k;
    if (i < p.n) {
      unsigned key = keys[i];
      h[k] = (i == 0) || key != prev; prev = key;
      ps[k] = pos[i];
      if (PHASE == 0) { x[k] = scan_input<A, OP>(p, ps[k], h[k]); if (p.gath) ((A*)p.gath)[i] = x[k]; }   // the only random gather of a
      else x[k] = p.gath ? ((const A*)p.gath)[i] : scan_input<A, OP>(p, ps[k], h[k]);
      if (!any) { tf = h[k]; tx = x[k]; any = true; }
      else { if (h[k]) tx = x[k]; else tx = ScanOp<A, OP>::apply(tx, x[k]); tf = tf | h[k]; }
      x[k] = tx;                               // thread-local inclusive scan
    } else { h[k] = true; ps[k] = 0; x[k] = A(0); }
  }
  if (!any) { tf = true; tx = A(0); }          // padding acts as a segment head: never merges
  bool pf, have_p, agg_f; A px, agg_x;
  block_seg_scan<A, OP>(tf, tx, pf, px, have_p, agg_f, agg_x);
  if (PHASE == 0) {
    if (threadIdx.x == 0) { tile_f[blockIdx.x] = agg_f; tile_x[blockIdx.x] = agg_x; }
    return;
  }
  // carry from preceding tiles applies until the first head flag in this tile
  bool have_c = blockIdx.x > 0 && carry_ok[blockIdx.x];
  A cx = have_c ? carry_x[blockIdx.x] : A(0);
  // full exclusive prefix for this thread = carry (+) in-tile prefix
  bool have = false; A pre = A(0);
  if (have_p) { pre = px; have = true; if (have_c && !pf) pre = ScanOp<A, OP>::apply(cx, px); }
  else if (have_c) { pre = cx; have = true; }
  bool open = have;                            // prefix still applies while no head has been seen in this thread
#pragma unroll
  for (int k = 0; k < SG_ITEMS; ++k) {
    long long i = base + k;
    if (i >= p.n) break;
    if (h[k]) open = false;
    A r = open ? ScanOp<A, OP>::apply(pre, x[k]) : x[k];
    if (p.stage && sel) { ScanParams q = p; q.out = p.stage; scan_store<A, OP>(q, i, r); }
    else scan_store<A, OP>(p, (long long)ps[k], r);
  }
}

// ---------------------------------------------------------------------------------------------
// Un-sort: out[pos[k]] = staged[k].  Done naively this is one random 8-byte store per element over the whole output (2 GB
// at N = 2^28): 18 of the 30 ms of an unsorted cumsum, 85 % of the store sectors missing L2 and paying a read-modify-write.
// Instead (pos, value) is partitioned by the top 8 bits of pos with one pass of the radix scatter (coalesced), and the
// partitioned records are then stored by CTAs that walk them IN ORDER: the CTAs resident at any time work on a few
// neighbouring output windows (8 MB each at N = 2^28), which stay in L2 until their sectors are complete.
// ---------------------------------------------------------------------------------------------
template <class V>
__global__ void __launch_bounds__(256) k_window_scatter(const unsigned* pos, const V* val, long long n, V* out, const int* run_flag) {
  if (run_flag && *run_flag == 0) return;
  // one CTA per 8192 records, started in blockIdx order: the CTAs resident together cover a few neighbouring windows.
  // (A persistent grid walking the chunks round robin measured 10.1 ms against 4.3: its CTAs drift apart and touch many
  //  windows at once -- DRAM traffic 19 GB against 11.6.)
  const long long base = (long long)blockIdx.x * 8192;
#pragma unroll 4
  for (int c = 0; c < 32; ++c) {
    const long long i = base + c * 256 + threadIdx.x;
    if (i < n) out[pos[i]] = val[i];
  }
}
template <class V>
inline void unsort_windowed(const unsigned* pos, const V* staged, long long n, V* out, unsigned* pos_part, V* val_part,
                            unsigned* hist, unsigned* tile_sums, const int* run_flag, cudaStream_t st) {
  int bits = 1; while (bits < 32 && (1ll << bits) < n) ++bits;
  const int shift = bits > 8 ? bits - 8 : 0;
  const int nb = rs_blocks_k<unsigned>(n);
  static std::atomic<unsigned long long> attr_done{0};
  ensure_dyn_smem(k_radix_scatter<unsigned, V>, (int)rs_scatter_smem<unsigned, V>(), attr_done);
  AGC_COUNT_LAUNCH(1), k_radix_hist<unsigned><<<tile_grid(nb, 4), RS_THREADS, 0, st>>>(pos, n, shift, hist, nb, run_flag);
  exclusive_scan_u32(hist, 256ll * nb, tile_sums, st, run_flag);
  AGC_COUNT_LAUNCH(1), k_radix_scatter<unsigned, V><<<tile_grid(nb, 2), RS_THREADS, rs_scatter_smem<unsigned, V>(), st>>>(pos, staged, pos_part, val_part, n, shift, hist, nb, run_flag);
  AGC_COUNT_LAUNCH(1), k_window_scatter<V><<<(unsigned)((n + 8191) / 8192), 256, 0, st>>>(pos_part, val_part, n, out, run_flag);
}

// ---- vectorised element access for the direct scan: 8 consecutive elements per thread as 128-bit loads ----
__device__ __forceinline__ bool raw8_ok(int dt) { return dt == AGC_DT_I64 || dt == AGC_DT_U64 || dt == AGC_DT_F64 || dt == AGC_DT_F32 || dt == AGC_DT_I32 || dt == AGC_DT_U32; }
__device__ __forceinline__ void load_raw8(const void* a, int dt, long long base, unsigned long long (&raw)[SG_ITEMS]) {
  if (dtype_bytes(dt) == 8) {
    const int4* q = (const int4*)((const unsigned long long*)a + base);
#pragma unroll
    for (int k = 0; k < SG_ITEMS / 2; ++k) {
      const int4 v = q[k];
      raw[2 * k] = ((unsigned long long)(unsigned)v.x) | ((unsigned long long)(unsigned)v.y << 32);
      raw[2 * k + 1] = ((unsigned long long)(unsigned)v.z) | ((unsigned long long)(unsigned)v.w << 32);
    }
  } else {
    const int4* q = (const int4*)((const unsigned*)a + base);
#pragma unroll
    for (int k = 0; k < SG_ITEMS / 4; ++k) {
      const int4 v = q[k];
      raw[4 * k] = (unsigned)v.x; raw[4 * k + 1] = (unsigned)v.y; raw[4 * k + 2] = (unsigned)v.z; raw[4 * k + 3] = (unsigned)v.w;
    }
  }
}
// same value semantics as scan_input, from the element's raw bits
template <class A, int OP>
__device__ __forceinline__ A scan_from_raw(const ScanParams& p, unsigned long long raw, bool head) {
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  const int dt = p.a_dtype;
  const long long vi = dt == AGC_DT_I32 ? (long long)(int)(unsigned)raw : (long long)raw;      // U32 / U64 / I64: raw as is
  if (OP == SC_SUM || OP == SC_PROD) {
    if (std::is_same<A, double>::value) {
      double v;
      if (dt == AGC_DT_F32) v = (double)__uint_as_float((unsigned)raw);
      else if (dt == AGC_DT_F64) v = __longlong_as_double((long long)raw);
      else if (p.a_u64) v = (double)raw;
      else v = (double)vi;
      if (nanskip && v != v) v = OP == SC_SUM ? 0.0 : 1.0;
      return (A)v;
    } else {
      return (A)vi;
    }
  } else {
    const long long ident = OP == SC_MAX ? AGC_EMPTY_MAX : AGC_EMPTY_MIN;
    const long long absorb = OP == SC_MAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX;
    if (p.a_float) {
      const double v = dt == AGC_DT_F32 ? (double)__uint_as_float((unsigned)raw) : __longlong_as_double((long long)raw);
      if (v != v) return (A)((nanskip || !head) ? ident : absorb);
      return (A)key_of(v);
    }
    return (A)(p.a_u64 ? key_of((unsigned long long)raw) : vi);
  }
}

// Warp-cooperative transposed access for 8-byte elements: a warp owns 256 consecutive elements (2 KB).  Memory is
// touched with fully coalesced 128-bit accesses (lane l -> chunk k*32 + l), the registers end up BLOCKED (thread t holds
// elements 8t .. 8t+7) as the thread-local scan needs.  The 2 KB staging buffer per warp is XOR-swizzled
// (chunk c lives at c ^ ((c >> 2) & 7)) so that both the striped and the blocked side are bank-conflict free.
// Per-thread 64-byte loads reach the same DRAM bytes but cost 4x the L1 wavefronts (ncu: L1TEX 98 % busy).
// step 1: the coalesced global loads of one 2 KB warp slice (lane l -> chunk k * 32 + l), nothing else -- callers issue the
// loads of labels AND values before they touch either (one memory round trip per tile instead of two)
__device__ __forceinline__ void warp_fetch8(const void* src_warp, int lane, int4 (&raw)[4]) {
  const int4* q = (const int4*)src_warp;
#pragma unroll
  for (int k = 0; k < 4; ++k) raw[k] = q[k * 32 + lane];
}
// step 2: through the XOR-swizzled staging buffer into BLOCKED registers (thread t holds elements 8t .. 8t+7)
__device__ __forceinline__ void warp_transpose8(const int4 (&raw)[4], int4* wbuf, int lane, unsigned long long (&out)[SG_ITEMS]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) { const int c = k * 32 + lane; wbuf[c ^ ((c >> 2) & 7)] = raw[k]; }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * lane + j;
    const int4 v = wbuf[c ^ ((c >> 2) & 7)];
    out[2 * j] = ((unsigned long long)(unsigned)v.x) | ((unsigned long long)(unsigned)v.y << 32);
    out[2 * j + 1] = ((unsigned long long)(unsigned)v.z) | ((unsigned long long)(unsigned)v.w << 32);
  }
  __syncwarp();
}
__device__ __forceinline__ void warp_load_blocked8(const void* src_warp, int4* wbuf, int lane, unsigned long long (&out)[SG_ITEMS]) {
  int4 raw[4];
  warp_fetch8(src_warp, lane, raw);
  warp_transpose8(raw, wbuf, lane, out);
}
__device__ __forceinline__ void warp_store_blocked8(void* dst_warp, int4* wbuf, int lane, const unsigned long long (&in)[SG_ITEMS]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * lane + j;
    wbuf[c ^ ((c >> 2) & 7)] = make_int4((int)(unsigned)in[2 * j], (int)(in[2 * j] >> 32), (int)(unsigned)in[2 * j + 1], (int)(in[2 * j + 1] >> 32));
  }
  __syncwarp();
  int4* q = (int4*)dst_warp;
#pragma unroll
  for (int k = 0; k < 4; ++k) { const int c = k * 32 + lane; q[c] = wbuf[c ^ ((c >> 2) & 7)]; }
  __syncwarp();
}


// ---------------------------------------------------------------------------------------------
// single-pass chained scan: tile descriptors + warp-parallel decoupled look-back
// ---------------------------------------------------------------------------------------------
// One 16-byte descriptor per tile, written and read as ONE 128-bit access: {meta, value bits}, meta = status (0 not
// there yet, 1 tile aggregate, 2 inclusive prefix) | 0x100 when the tile holds a segment head (its value then is the
// running value of the tile's LAST segment and nothing before the tile matters).  Warp 0 of tile t looks at the 32
// tiles before it at once (lane l -> tile t - 1 - l), waits until they have published anything, folds the values up
// to the nearest "terminator" (a prefix, or an aggregate with a head) with a bounded shuffle tree, and moves on by 32
// tiles only if the window held none.  Tiles run in blockIdx order (tile 0 publishes its prefix at once), so every
// descriptor a warp waits for belongs to a CTA that is already running or done.
__device__ __forceinline__ void desc_store(unsigned long long* d, unsigned long long meta, unsigned long long xbits) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" :: "l"(d), "l"(meta), "l"(xbits) : "memory");
}
__device__ __forceinline__ void desc_load(const unsigned long long* d, unsigned long long& meta, unsigned long long& xbits) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(meta), "=l"(xbits) : "l"(d) : "memory");
}
// the same with GPU scope instead of the system scope `volatile` implies (experiment: AGC_DESC_GPU=1)
__device__ __forceinline__ void desc_store_gpu(unsigned long long* d, unsigned long long meta, unsigned long long xbits) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(d), "l"(meta), "l"(xbits) : "memory");
}
__device__ __forceinline__ void desc_load_gpu(const unsigned long long* d, unsigned long long& meta, unsigned long long& xbits) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(meta), "=l"(xbits) : "l"(d) : "memory");
}
template <class A> __device__ __forceinline__ unsigned long long scan_bits(A x);
template <> __device__ __forceinline__ unsigned long long scan_bits<long long>(long long x) { return (unsigned long long)x; }
template <> __device__ __forceinline__ unsigned long long scan_bits<double>(double x) { return (unsigned long long)__double_as_longlong(x); }
template <class A> __device__ __forceinline__ A scan_unbits(unsigned long long b);
template <> __device__ __forceinline__ long long scan_unbits<long long>(unsigned long long b) { return (long long)b; }
template <> __device__ __forceinline__ double scan_unbits<double>(unsigned long long b) { return __longlong_as_double((long long)b); }

// called by the 32 threads of warp 0; returns the carry into tile t (t > 0): the running value at the end of tile t - 1.
// A round fetches LB_W windows of 32 descriptors at once (lane l: tiles t-1-l, t-33-l, ...).  Measured at N = 2^28
// (cumsum, sorted int64 labels): LB_W = 1 2.24 ms, LB_W = 4 2.34 ms -- a wider round waits for more predecessors to
// publish and gains nothing, the look-back is bound by WHEN the aggregates appear, not by how many it can read at once.
constexpr int LB_W = 1;
template <class A, int OP>
__device__ __forceinline__ A chain_lookback(const unsigned long long* desc, int t, int lane, unsigned sleep_ns = 0) {
  A acc = A(0); bool have = false;
  int look = t - 1 - lane;
  for (;;) {
    unsigned long long m[LB_W], xb[LB_W];
#pragma unroll
    for (int w = 0; w < LB_W; ++w) {                           // all windows of the round in flight together
      m[w] = 2ull; xb[w] = 0ull;                               // tiles before tile 0 never matter (tile 0 is a terminator)
      if (look - 32 * w >= 0) desc_load(desc + 2 * (long long)(look - 32 * w), m[w], xb[w]);
    }
    bool done = false;
#pragma unroll
    for (int w = 0; w < LB_W; ++w) {
      if (done) break;
      const int lk = look - 32 * w;
      while (lk >= 0 && (m[w] & 3ull) == 0ull) {               // not published yet: wait for it (backing off keeps the pollers
        if (sleep_ns) __nanosleep(sleep_ns);                   // from fighting the publisher for the descriptor's L2 line)
        desc_load(desc + 2 * (long long)lk, m[w], xb[w]);
      }
      const bool term = (m[w] & 3ull) == 2ull || (m[w] & 0x100ull) != 0ull;
      const unsigned tb = __ballot_sync(0xffffffffu, term);
      const int first = tb ? __ffs((int)tb) - 1 : 32;          // nearest terminator in this window
      A x = scan_unbits<A>(xb[w]);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {                       // lane i ends with the fold of lanes [i, first], farthest first
        const A y = __shfl_down_sync(0xffffffffu, x, o);
        if (lane + o <= first && lane + o < 32) x = ScanOp<A, OP>::apply(y, x);
      }
      const A win = __shfl_sync(0xffffffffu, x, 0);
      acc = have ? ScanOp<A, OP>::apply(win, acc) : win; have = true;
      if (tb) done = true;
    }
    if (done) break;
    look -= 32 * LB_W;
  }
  return acc;
}

// Direct variant for flat int32/int64 labels that are already non-decreasing: no keys, no positions, no sort.
// PHASE 0 reads (label, value) once for the tile aggregates and, on the way, verifies sortedness and label
// bounds; PHASE 1 re-reads them, applies the carry and writes out[i] -- 40 B/element instead of the 64+ of the
// keys path.  If PHASE 0 finds an inversion it raises *unsorted and PHASE 1 returns at once; the sort-based path,
// gated on the same flag, then does the work.
template <class LT> __device__ __forceinline__ void load_labels8(const void* labels, long long base, long long n, long long (&g)[SG_ITEMS]);
template <> __device__ __forceinline__ void load_labels8<long long>(const void* labels, long long base, long long n, long long (&g)[SG_ITEMS]) {
  if (base + SG_ITEMS <= n) {
    const int4* q = (const int4*)((const long long*)labels + base);
#pragma unroll
    for (int k = 0; k < SG_ITEMS / 2; ++k) {
      int4 v = q[k];
      g[2 * k] = ((long long)(unsigned)v.x) | ((long long)v.y << 32); g[2 * k + 1] = ((long long)(unsigned)v.z) | ((long long)v.w << 32);
    }
  } else {
#pragma unroll
    for (int k = 0; k < SG_ITEMS; ++k) g[k] = base + k < n ? ((const long long*)labels)[base + k] : 0;
  }
}
template <> __device__ __forceinline__ void load_labels8<int>(const void* labels, long long base, long long n, long long (&g)[SG_ITEMS]) {
  if (base + SG_ITEMS <= n) {
    const int4* q = (const int4*)((const int*)labels + base);
#pragma unroll
    for (int k = 0; k < SG_ITEMS / 4; ++k) { int4 v = q[k]; g[4 * k] = v.x; g[4 * k + 1] = v.y; g[4 * k + 2] = v.z; g[4 * k + 3] = v.w; }
  } else {
#pragma unroll
    for (int k = 0; k < SG_ITEMS; ++k) g[k] = base + k < n ? ((const int*)labels)[base + k] : 0;
  }
}

template <class A, int OP, class LT, int PHASE>
__global__ void __launch_bounds__(SG_THREADS) k_seg_scan_direct(const ScanParams p, unsigned char* tile_f, A* tile_x,
                                                                const unsigned char* carry_ok, const A* carry_x) {
  // PHASE 1: nothing to do when the labels are not sorted; PHASE 0: another tile already found an inversion.
  // (block-uniform decision: every thread takes the same branch around the barriers below)
  // PHASE 2 (single pass, chained): tile_x carries the tile descriptors; a tile that gives up still publishes one
  const int tile_id = (int)blockIdx.x + p.tile0;
  if (__syncthreads_or(*(volatile const int*)p.unsorted)) {
    if (PHASE == 2 && threadIdx.x == 0) desc_store((unsigned long long*)tile_x + 2 * (long long)tile_id, 2ull | 0x100ull, 0ull);
    return;
  }
  const long long base = (long long)tile_id * SG_TILE + (long long)threadIdx.x * SG_ITEMS;
  A x[SG_ITEMS]; bool h[SG_ITEMS];
  long long g[SG_ITEMS];
  __shared__ int4 s_stage[SG_THREADS / 32][128];            // 2 KB per warp: transposed 8-byte accesses
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const long long wbase = (long long)tile_id * SG_TILE + (long long)wrp * (32 * SG_ITEMS);
  const bool wfull = wbase + 32 * SG_ITEMS <= p.n;          // warp-uniform
  const bool tr_lab = wfull && sizeof(LT) == 8;
  const bool tr_val = wfull && dtype_bytes(p.a_dtype) == 8 && !((uintptr_t)p.a & 15);
  // both transposed streams: all global loads first, then the two trips through the staging buffer
  int4 raw_lab[4], raw_val[4];
  if (tr_lab) warp_fetch8((const long long*)p.labels + wbase, lane, raw_lab);
  if (tr_val) warp_fetch8((const unsigned long long*)p.a + wbase, lane, raw_val);
  if (tr_lab) {
    unsigned long long gl[SG_ITEMS];
    warp_transpose8(raw_lab, s_stage[wrp], lane, gl);
#pragma unroll
    for (int k = 0; k < SG_ITEMS; ++k) g[k] = (long long)gl[k];
  } else load_labels8<LT>(p.labels, base, p.n, g);
  long long prev = 0;
  const long long left = __shfl_up_sync(0xffffffffu, g[SG_ITEMS - 1], 1);      // the previous thread's last label
  if (wfull && lane > 0) prev = left;
  else if (base > 0 && base - 1 < p.n) prev = sizeof(LT) == 8 ? ((const long long*)p.labels)[base - 1] : (long long)((const int*)p.labels)[base - 1];
  bool tf = false; A tx = A(0); bool any = false, uns = false, bad = false;
  const bool vec = base + SG_ITEMS <= p.n && raw8_ok(p.a_dtype) && !((uintptr_t)p.a & 15);
  unsigned long long raw[SG_ITEMS];
  if (tr_val) warp_transpose8(raw_val, s_stage[wrp], lane, raw);
  else if (vec) load_raw8(p.a, p.a_dtype, base, raw);
#pragma unroll
  for (int k = 0; k < SG_ITEMS; ++k) {
    const long long i = base + k;
    if (i < p.n) {
      h[k] = (i == 0) || g[k] != prev;
      if (PHASE != 1) { uns |= i > 0 && g[k] < prev; bad |= g[k] < 0 || g[k] >= p.size; }
      prev = g[k];
      x[k] = vec ? scan_from_raw<A, OP>(p, raw[k], h[k]) : scan_input<A, OP>(p, (unsigned)i, h[k]);
      if (!any) { tf = h[k]; tx = x[k]; any = true; }
      else { if (h[k]) tx = x[k]; else tx = ScanOp<A, OP>::apply(tx, x[k]); tf = tf | h[k]; }
      x[k] = tx;                               // thread-local inclusive scan
    } else { h[k] = true; x[k] = A(0); }
  }
  if (PHASE != 1) {
    if (uns) *p.unsorted_w = 1;
    if (bad) p.status[AGC_ST_BADLABEL] = 1;
  }
  if (!any) { tf = true; tx = A(0); }
  bool pf, have_p, agg_f; A px, agg_x;
  block_seg_scan<A, OP>(tf, tx, pf, px, have_p, agg_f, agg_x);
  if (PHASE == 0) {
    if (threadIdx.x == 0) { tile_f[tile_id] = agg_f; tile_x[tile_id] = agg_x; }
    return;
  }
  bool have_c; A cx;
  if (PHASE == 2) {
    __shared__ unsigned long long s_carry;
    if (threadIdx.x < 32) {
      unsigned long long* desc = (unsigned long long*)tile_x;
      const int t = tile_id;
      // the aggregate first (tile 0: it already is the inclusive prefix), then the look-back, then the prefix
      if (lane == 0) desc_store(desc + 2 * (long long)t, (t == 0 ? 2ull : 1ull) | (agg_f ? 0x100ull : 0ull), scan_bits<A>(agg_x));
      A c = A(0);
      if (t > 0) {
        c = chain_lookback<A, OP>(desc, t, lane, (unsigned)p.lb_sleep & 0xffffu);
        if (lane == 0) desc_store(desc + 2 * (long long)t, 2ull | 0x100ull, scan_bits<A>(agg_f ? agg_x : ScanOp<A, OP>::apply(c, agg_x)));
      }
      if (lane == 0) s_carry = scan_bits<A>(c);
    }
    __syncthreads();
    have_c = tile_id > 0; cx = scan_unbits<A>(s_carry);
  } else {
    have_c = tile_id > 0 && carry_ok[tile_id];
    cx = have_c ? carry_x[tile_id] : A(0);
  }
  bool have = false; A pre = A(0);
  if (have_p) { pre = px; have = true; if (have_c && !pf) pre = ScanOp<A, OP>::apply(cx, px); }
  else if (have_c) { pre = cx; have = true; }
  bool open = have;
  // cumsum / cumprod into a 64-bit output of the accumulator's own class: results leave as 128-bit stores
  // integer running max / min keep the input dtype: for 64-bit integers the key IS the value (u64: sign bit flipped back)
  const bool mm64 = (OP == SC_MAX || OP == SC_MIN) && !p.a_float && (p.out_dtype == AGC_DT_I64 || p.out_dtype == AGC_DT_U64);
  const bool vst = base + SG_ITEMS <= p.n && !((uintptr_t)p.out & 15) &&
                   (mm64 || ((OP == SC_SUM || OP == SC_PROD) &&
                             (std::is_same<A, double>::value ? p.out_dtype == AGC_DT_F64 : (p.out_dtype == AGC_DT_I64 || p.out_dtype == AGC_DT_U64))));
  const unsigned long long keyflip = (mm64 && p.a_u64) ? 0x8000000000000000ull : 0ull;
  A res[SG_ITEMS];
#pragma unroll
  for (int k = 0; k < SG_ITEMS; ++k) {
    const long long i = base + k;
    if (i >= p.n) break;
    if (h[k]) open = false;
    A r = open ? ScanOp<A, OP>::apply(pre, x[k]) : x[k];
    res[k] = r;
    if (!vst) scan_store<A, OP>(p, i, r);
  }
  if (vst && wfull) {                                       // coalesced 128-bit stores through the warp's staging buffer
    unsigned long long ob[SG_ITEMS];
#pragma unroll
    for (int k = 0; k < SG_ITEMS; ++k)
      ob[k] = std::is_same<A, double>::value ? (unsigned long long)__double_as_longlong((double)res[k]) : ((unsigned long long)(long long)res[k] ^ keyflip);
    warp_store_blocked8((A*)p.out + wbase, s_stage[wrp], lane, ob);
  } else if (vst) {
    int4* q = (int4*)((A*)p.out + base);
#pragma unroll
    for (int k = 0; k < SG_ITEMS / 2; ++k) {
      const long long b0 = std::is_same<A, double>::value ? __double_as_longlong((double)res[2 * k]) : (long long)((unsigned long long)(long long)res[2 * k] ^ keyflip);
      const long long b1 = std::is_same<A, double>::value ? __double_as_longlong((double)res[2 * k + 1]) : (long long)((unsigned long long)(long long)res[2 * k + 1] ^ keyflip);
      q[k] = make_int4((int)(unsigned)b0, (int)(b0 >> 32), (int)(unsigned)b1, (int)(b1 >> 32));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_seg_scan_lean: the chained scan for the all-64-bit case (int64 labels, int64 / uint64 / float64 values, result of
// the accumulator's own width, full tiles) with everything the general kernel decides at run time -- value dtype,
// nan-skip, output conversion, ragged edges -- fixed at compile time.  The general PHASE-2 kernel executes 142 thread
// instructions per element (ncu, r02e: ISETP / MOV / SEL / BRA for dtype switches and bounds it re-derives per element)
// at 46 % issue utilisation: its 2.2 ms are as much instruction count as look-back latency.  VK: 0 int64, 1 uint64,
// 2 float64 values.  Tiles past the last full one go to k_seg_scan_direct<.., 2> with tile0 set; the descriptors are shared.
// (A persistent variant of this kernel with the next tile's loads in flight -- 128 registers, 2 CTAs per SM -- measured
// 3.08 ms against 2.24: with two resident CTAs nothing runs while one waits for its look-back.)
// ---------------------------------------------------------------------------------------------
template <class A, int OP, int VK, bool NANSKIP, int OCC>
__global__ void __launch_bounds__(SG_THREADS, OCC) k_seg_scan_lean(const ScanParams p, unsigned long long* desc) {
  __shared__ int4 s_stage[SG_THREADS / 32][128];
  __shared__ unsigned long long s_carry;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int tile = blockIdx.x;
  if (__syncthreads_or(*(volatile const int*)p.unsorted)) {     // somebody found an inversion: the sort-based path redoes everything
    if (threadIdx.x == 0) desc_store(desc + 2 * (long long)tile, 2ull | 0x100ull, 0ull);
    return;
  }
  const long long base = (long long)tile * SG_TILE + (long long)threadIdx.x * SG_ITEMS;
  const long long wbase = (long long)tile * SG_TILE + (long long)wrp * (32 * SG_ITEMS);
  int4 raw_lab[4], raw_val[4];
  warp_fetch8((const long long*)p.labels + wbase, lane, raw_lab);
  warp_fetch8((const unsigned long long*)p.a + wbase, lane, raw_val);
  unsigned long long gl[SG_ITEMS], raw[SG_ITEMS];
  warp_transpose8(raw_lab, s_stage[wrp], lane, gl);
  warp_transpose8(raw_val, s_stage[wrp], lane, raw);
  long long prev = 0;
  const long long left = __shfl_up_sync(0xffffffffu, (long long)gl[SG_ITEMS - 1], 1);
  if (lane > 0) prev = left;
  else if (base > 0) prev = ((const long long*)p.labels)[base - 1];
  A x[SG_ITEMS]; unsigned hmask = 0;                           // bit k: element k opens a segment
  A tx = A(0); bool uns = false, bad = false;
  const unsigned long long keyflip = (VK == 1 && (OP == SC_MAX || OP == SC_MIN)) ? 0x8000000000000000ull : 0ull;
#pragma unroll
  for (int k = 0; k < SG_ITEMS; ++k) {
    const long long g = (long long)gl[k];
    const bool head = (k == 0 && base == 0) || g != prev;
    uns |= g < prev && !(k == 0 && base == 0);
    bad |= (unsigned long long)g >= (unsigned long long)p.size;
    prev = g;
    A v;
    if (VK == 2) { double d = __longlong_as_double((long long)raw[k]); if (NANSKIP && d != d) d = OP == SC_SUM ? 0.0 : 1.0; v = (A)d; }
    else v = (A)(long long)(raw[k] ^ keyflip);                 // int64 as is; uint64 max / min: order-preserving key
    tx = (k == 0 || head) ? v : ScanOp<A, OP>::apply(tx, v);
    hmask |= head ? (1u << k) : 0u;
    x[k] = tx;                                                 // thread-local inclusive scan
  }
  if (uns) *p.unsorted_w = 1;
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  bool pf, have_p, agg_f; A px, agg_x;
  block_seg_scan<A, OP>(hmask != 0u, tx, pf, px, have_p, agg_f, agg_x);
  if (threadIdx.x < 32) {
    if (lane == 0) desc_store(desc + 2 * (long long)tile, (tile == 0 ? 2ull : 1ull) | (agg_f ? 0x100ull : 0ull), scan_bits<A>(agg_x));
    A c = A(0);
    if (tile > 0) {
      c = chain_lookback<A, OP>(desc, tile, lane, (unsigned)p.lb_sleep & 0xffffu);
      if (lane == 0) desc_store(desc + 2 * (long long)tile, 2ull | 0x100ull, scan_bits<A>(agg_f ? agg_x : ScanOp<A, OP>::apply(c, agg_x)));
    }
    if (lane == 0) s_carry = scan_bits<A>(c);
  }
  __syncthreads();
  const bool have_c = tile > 0; const A cx = scan_unbits<A>(s_carry);
  bool have = false; A pre = A(0);
  if (have_p) { pre = px; have = true; if (have_c && !pf) pre = ScanOp<A, OP>::apply(cx, px); }
  else if (have_c) { pre = cx; have = true; }
  // the prefix applies to the elements before this thread's first segment head
  const int nopen = !have ? 0 : (hmask ? __ffs((int)hmask) - 1 : SG_ITEMS);
  unsigned long long ob[SG_ITEMS];
#pragma unroll
  for (int k = 0; k < SG_ITEMS; ++k) {
    const A r = k < nopen ? ScanOp<A, OP>::apply(pre, x[k]) : x[k];
    ob[k] = (std::is_same<A, double>::value ? (unsigned long long)__double_as_longlong((double)r) : (unsigned long long)(long long)r) ^ keyflip;
  }
  warp_store_blocked8((A*)p.out + wbase, s_stage[wrp], lane, ob);
}

// (Measured and removed: the same kernel with the WARP as the chained unit -- a 256-element tile and a descriptor per warp, no
//  CTA barrier at all.  1.68 ms at 5 CTAs per SM, 1.76 at 6, against 1.55 here: eight times the descriptors, a dependent global
//  load for the label left of every warp tile, and the duty cycle of the loads no better.  profiles/r02i_scan_variants.jsonl)
// ---------------------------------------------------------------------------------------------
// The chained scan as a PERSISTENT kernel fed by the TMA (k_seg_scan_ws below; these are its building blocks).  What the
// one-tile-per-CTA kernels above cannot fix is their duty cycle: a tile has its loads in flight for ~2.5 us of an 8 us life,
// and even the plain two-pass kernels (no chain at all) stream at 4.5-5.3 TB/s.  A persistent CTA keeps the loads of the next
// SP_STAGES tiles in flight with cp.async.bulk.tensor into a ring in shared memory (complete_tx on one mbarrier per stage:
// 16 KB of labels, 16 KB of values, and the 16 bytes before the tile for the label left of it) -- the bytes in flight no longer
// depend on what the threads are doing -- and the results leave through shared memory with bulk tensor stores.
// 3 stages * 32 KB + 16 KB of results = 112 KB per CTA, two CTAs per SM.
// (History, profiles/r02i_scan_variants.jsonl: the first persistent kernel kept the CTA-wide structure -- block scan behind two
//  barriers, warp 0's look-back with seven warps waiting, one 16 KB store per tile: cumsum 1.28 ms, cummax 1.43-1.57;
//  k_seg_scan_ws replaced it.)
// ---------------------------------------------------------------------------------------------
constexpr int SP_STAGES = 3;
constexpr int SP_IN_BYTES = SG_TILE * 8;                         // one array of one tile
__device__ __forceinline__ unsigned sp_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sp_mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void sp_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void sp_bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void sp_bulk_store(void* dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void sp_store_read_done() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void sp_store_done() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sp_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sp_mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void sp_tensor_load(unsigned dst, const void* tmap, int row, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(dst), "l"(tmap), "r"(0), "r"(row), "r"(bar) : "memory");
}
__device__ __forceinline__ void sp_tensor_store(const void* tmap, int row, unsigned src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" :: "l"(tmap), "r"(0), "r"(row), "r"(src) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// The tile buffers are filled by the TMA as boxes of 256 rows x 64 bytes with the 64-byte swizzle: the 16-byte unit k of
// row r lives at unit k ^ ((r >> 1) & 3) of the row (address bits 4-5 ^= bits 7-8; the buffers are 512-byte aligned).
// Row t is exactly thread t's 8 consecutive elements, and the XOR makes the quarter-warps' 128-bit accesses conflict
// free -- the registers come out in element order with no shuffling at all.  (A linear buffer read in a rotated order and
// rotated back in registers cost 200 of the kernel's 674 instructions per warp and tile.)
__device__ __forceinline__ void sp_read8(const int4* buf, int t, unsigned long long (&out)[SG_ITEMS]) {
  const int r = (t >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int4 u = buf[4 * t + (j ^ r)];
    out[2 * j] = ((unsigned long long)(unsigned)u.x) | ((unsigned long long)(unsigned)u.y << 32);
    out[2 * j + 1] = ((unsigned long long)(unsigned)u.z) | ((unsigned long long)(unsigned)u.w << 32);
  }
}
__device__ __forceinline__ void sp_write8(int4* buf, int t, const unsigned long long (&in)[SG_ITEMS]) {
  const int r = (t >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    buf[4 * t + (j ^ r)] = make_int4((int)(unsigned)in[2 * j], (int)(in[2 * j] >> 32), (int)(unsigned)in[2 * j + 1], (int)(in[2 * j + 1] >> 32));
}
// look-back with the first window's descriptors already requested (m, xb)
struct SpNoHook { __device__ __forceinline__ void operator()() const {} };
// `on_wait` runs (all 32 lanes, converged) every time a poll came back without what the look-back needs
template <class A, int OP, class F = SpNoHook>
__device__ __forceinline__ A chain_lookback_pre(const unsigned long long* desc, int t, int lane, unsigned long long m, unsigned long long xb,
                                                unsigned sleep_ns = 0, bool gpu_scope = false, F on_wait = F()) {
  A acc = A(0); bool have = false;
  int look = t - 1 - lane;
  for (;;) {
    int first;
    for (;;) {                                                 // wait only for the descriptors up to the nearest terminator seen so far
      const bool pub = (m & 3ull) != 0ull;
      const bool term = pub && ((m & 3ull) == 2ull || (m & 0x100ull) != 0ull);
      const unsigned pb = __ballot_sync(0xffffffffu, pub), tb = __ballot_sync(0xffffffffu, term);
      first = tb ? __ffs((int)tb) - 1 : 32;
      const unsigned need = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);
      if ((pb & need) == need) break;
      on_wait();
      if (sleep_ns) __nanosleep(sleep_ns);
      if (!pub && lane <= first) { if (gpu_scope) desc_load_gpu(desc + 2 * (long long)look, m, xb); else desc_load(desc + 2 * (long long)look, m, xb); }
    }
    A x = scan_unbits<A>(xb);
    if (first > 0) {                                           // (the tile right before is a terminator: its value is the carry)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {                       // lane i ends with the fold of lanes [i, first], farthest first
        const A y = __shfl_down_sync(0xffffffffu, x, o);
        if (lane + o <= first && lane + o < 32) x = ScanOp<A, OP>::apply(y, x);
      }
    }
    const A win = __shfl_sync(0xffffffffu, x, 0);
    acc = have ? ScanOp<A, OP>::apply(win, acc) : win; have = true;
    if (first < 32) break;
    look -= 32;
    m = 2ull; xb = 0ull;                                       // tiles before tile 0 never matter
    if (look >= 0) { if (gpu_scope) desc_load_gpu(desc + 2 * (long long)look, m, xb); else desc_load(desc + 2 * (long long)look, m, xb); }
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// k_seg_scan_ws: the persistent chained scan without a single CTA-wide barrier -- the block-level part of the scan, the
// look-back and the loads belong to a NINTH warp.  ncu on its predecessor with the CTA-wide structure (r2q): the memory
// pipeline never runs dry (the mbarrier wait succeeds at the first try, always); what bounds that kernel is the serial path
// of a CTA through one tile -- four barriers, the warp-0 look-back with seven warps waiting (28 % of all stall samples), the
// serial folds of the block scan -- 2.4 us per 2048 elements, two CTAs per SM.  Here
//   * the 8 data warps read their 256 elements, scan them with shuffles, hand (flag, value) of the warp to the ninth
//     warp through shared memory + an mbarrier arrive, and go straight on to the NEXT tile; one tile later they pick up
//     their warp's exclusive prefix (carry included), finish the owed results and store their own 2 KB slice with a
//     32-row tensor store -- warps drift freely, nothing waits for the slowest warp;
//   * the ninth warp waits for the 8 arrivals, scans the 8 warp aggregates, refills the stage the tile occupied (its
//     readers have all arrived), publishes the tile's descriptor, runs the look-back (requested before the wait), publishes
//     the prefix and posts the 8 warp prefixes.
// Slots are double-buffered by tile parity; a data warp is at most one tile ahead of the ninth warp.  Launched cooperatively
// (all CTAs resident, see the prologue).
// ---------------------------------------------------------------------------------------------
constexpr int WS_THREADS = SG_THREADS + 32;
constexpr int WS_SMEM = 115712;                                   // two CTAs per SM: (228 KB - 2 * 1 KB reserved) / 2; 496 B of small state + alignment + 7 tile buffers
__device__ __forceinline__ bool sp_mbar_test(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void sp_mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }

// LB: bytes per label (8: int64, 4: int32 -- rows of 32 bytes with the 32-byte swizzle: unit k of row t sits at k ^ ((t >> 2) & 1))
template <class A, int OP, int VK, bool NANSKIP, int LB>
__global__ void __launch_bounds__(WS_THREADS, 2) k_seg_scan_ws(const ScanParams p, unsigned long long* desc, int nfull,
                                                               const __grid_constant__ CUtensorMap tm_lab, const __grid_constant__ CUtensorMap tm_val,
                                                               const __grid_constant__ CUtensorMap tm_out32) {
  extern __shared__ __align__(16) unsigned char sp_dyn[];
  // small state first (496 bytes), then the tile buffers at the next 512-byte boundary (the swizzle is a function of the address)
  int4* s_left = (int4*)sp_dyn;                                          // [stage]: the 16 bytes before the tile's labels
  unsigned long long* s_bar = (unsigned long long*)(s_left + SP_STAGES); // full[3], agg_ready[2], prefix_ready[2]
  unsigned long long* s_aggx = s_bar + 8;                                // [2][8]
  unsigned long long* s_prex = s_aggx + 16;                              // [2][8]
  unsigned* s_aggf = (unsigned*)(s_prex + 16);                           // [2][8]
  unsigned* s_preh = s_aggf + 16;                                        // [2][8]: does warp w have a prefix at all
  unsigned char* sp_raw = sp_dyn + ((sp_saddr(sp_dyn) + 496u + 511u) & ~511u) - sp_saddr(sp_dyn);
  int4* s_in = (int4*)sp_raw;                                            // [stage][labels | values][1024 units]
  int4* s_out = (int4*)(sp_raw + SP_STAGES * 2 * SP_IN_BYTES);           // [warp][128 units]
  const int t = threadIdx.x, lane = t & 31, wrp = t >> 5;
  const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
  const unsigned long long keyflip = (VK == 1 && (OP == SC_MAX || OP == SC_MIN)) ? 0x8000000000000000ull : 0ull;
  const unsigned bar_full = sp_saddr(s_bar), bar_agg = sp_saddr(s_bar + 3), bar_pre = sp_saddr(s_bar + 5);
  // the sampled pre-check already found an inversion: this word is written by that kernel only, so every thread of every CTA
  // reads the same value (the run-time flag p.unsorted may come up while the kernel runs -- leaving on it would strand the
  // CTAs that wait for this one's tiles)
  if (p.unsorted[1]) return;
  const unsigned sleep_ns = (unsigned)p.lb_sleep & 0xffffu; const bool gpu_scope = (p.lb_sleep >> 30) & 1;
  auto issue = [&](int s, int tile) {
    const unsigned bar = bar_full + 8 * s;
    sp_expect_tx(bar, SG_TILE * LB + SP_IN_BYTES + (tile > 0 ? 16 : 0));
    sp_tensor_load(sp_saddr(s_in + (2 * s) * (SG_TILE / 2)), &tm_lab, tile * SG_THREADS, bar);
    sp_tensor_load(sp_saddr(s_in + (2 * s + 1) * (SG_TILE / 2)), &tm_val, tile * SG_THREADS, bar);
    if (tile > 0) sp_bulk_load(sp_saddr(s_left + s), (const char*)p.labels + ((long long)tile * SG_TILE) * LB - 16, 16, bar);
  };
  if (t == SG_THREADS) {                                                 // lane 0 of the ninth warp sets everything up
#pragma unroll
    for (int s = 0; s < SP_STAGES; ++s) sp_mbar_init(bar_full + 8 * s, 1);
    sp_mbar_init(bar_agg, SG_THREADS / 32); sp_mbar_init(bar_agg + 8, SG_THREADS / 32);
    sp_mbar_init(bar_pre, 1); sp_mbar_init(bar_pre + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sp_fence_async();
    // Tiles go round robin: iteration `it` of CTA b owns tile it * gridDim + b, so that neighbouring tiles are worked on at the
    // same time by neighbouring CTAs.  The launch is COOPERATIVE: the runtime guarantees that every CTA of the grid is resident,
    // so every tile a look-back waits for is being worked on.  (Measured alternatives: tickets from a counter, three taken
    // back to back at the start -- CTA 0 held tiles 0, 1, 2 and CTA 1 waited two iterations for tile 2; tickets taken one
    // dependent atomic after the other -- 1.84 ms against 1.31: with tickets drawn three tiles ahead the order in which
    // tiles are WORKED ON drifts away from the order in which they were drawn, and everybody waits for the oldest one.)
#pragma unroll
    for (int s = 0; s < SP_STAGES; ++s) { const int tile = s * grid + cta; if (tile < nfull) issue(s, tile); }
  }
  __syncthreads();

  if (wrp == SG_THREADS / 32) {
    // ---------------- the ninth warp: block-level scan, descriptors, look-back, loads ----------------
    // Publishing a tile's aggregate must not wait for the look-back of the tile before it: the neighbour's look-back waits
    // for exactly that aggregate, and a serial "look back, then publish the next" turned every missed poll into a delay
    // handed down the grid (4 polls per tile, 2.5 us per iteration).  So while a look-back polls, the warp keeps checking
    // whether the data warps have delivered the NEXT tile and publishes it at once (one tile ahead at most: the data
    // warps wait for the prefix of tile i - 1 before they start tile i + 1).
    struct Pub { bool fe, agg_f; A xe, agg_x; unsigned long long lm, lxb; };
    auto publish = [&](int it, Pub& P) {                                 // the 8 warps of iteration `it` have arrived
      const int tile = it * grid + cta, s = it % SP_STAGES, slot = it & 1;
      bool f = true; A x = A(0);
      if (lane < SG_THREADS / 32) { f = s_aggf[slot * 8 + lane] != 0u; x = scan_unbits<A>(s_aggx[slot * 8 + lane]); }
      bool fi = f; A xi = x;                                             // inclusive over warps 0 .. lane
#pragma unroll
      for (int o = 1; o < SG_THREADS / 32; o <<= 1) {
        const bool ft = __shfl_up_sync(0xffffffffu, (int)fi, o);
        const A xt = __shfl_up_sync(0xffffffffu, xi, o);
        if (lane >= o) { if (!fi) xi = ScanOp<A, OP>::apply(xt, xi); fi = fi | ft; }
      }
      P.fe = __shfl_up_sync(0xffffffffu, (int)fi, 1); P.xe = __shfl_up_sync(0xffffffffu, xi, 1);
      P.agg_f = __shfl_sync(0xffffffffu, (int)fi, SG_THREADS / 32 - 1); P.agg_x = __shfl_sync(0xffffffffu, xi, SG_THREADS / 32 - 1);
      if (lane == 0) {
        if (gpu_scope) desc_store_gpu(desc + 2 * (long long)tile, (tile == 0 ? 2ull : 1ull) | (P.agg_f ? 0x100ull : 0ull), scan_bits<A>(P.agg_x));
        else desc_store(desc + 2 * (long long)tile, (tile == 0 ? 2ull : 1ull) | (P.agg_f ? 0x100ull : 0ull), scan_bits<A>(P.agg_x));
      }
      P.lm = 2ull; P.lxb = 0ull;
      if (tile - 1 - lane >= 0) { if (gpu_scope) desc_load_gpu(desc + 2 * (long long)(tile - 1 - lane), P.lm, P.lxb); else desc_load(desc + 2 * (long long)(tile - 1 - lane), P.lm, P.lxb); }
      if (lane == 0) {
        const int next = (it + SP_STAGES) * grid + cta;                  // every reader of stage s has arrived
        if (next < nfull) issue(s, next);
      }
    };
    Pub cur, nxt;
    bool have_nxt = false;
    for (int it = 0;; ++it) {
      const int tile = it * grid + cta;
      if (tile >= nfull) break;
      const int slot = it & 1;
      if (have_nxt) { cur = nxt; have_nxt = false; }
      else { sp_mbar_wait(bar_agg + 8 * slot, (unsigned)(it >> 1) & 1u); publish(it, cur); }
      A c = A(0);
      const bool have_c = tile > 0;
      if (have_c) {
        auto ahead = [&]() {
          const int nit = it + 1;
          if (have_nxt || nit * grid + cta >= nfull) return;
          const bool ok = sp_mbar_test(bar_agg + 8 * (nit & 1), (unsigned)(nit >> 1) & 1u);
          if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
            if (!ok) sp_mbar_wait(bar_agg + 8 * (nit & 1), (unsigned)(nit >> 1) & 1u);   // (lane 0 saw it complete: immediate)
            publish(nit, nxt); have_nxt = true;
          }
        };
        c = chain_lookback_pre<A, OP>(desc, tile, lane, cur.lm, cur.lxb, sleep_ns, gpu_scope, ahead);
        if (lane == 0) {
          const unsigned long long incl = scan_bits<A>(cur.agg_f ? cur.agg_x : ScanOp<A, OP>::apply(c, cur.agg_x));
          if (gpu_scope) desc_store_gpu(desc + 2 * (long long)tile, 2ull | 0x100ull, incl); else desc_store(desc + 2 * (long long)tile, 2ull | 0x100ull, incl);
        }
      }
      if (lane < SG_THREADS / 32) {                                      // warp `lane`: exclusive prefix inside the tile, carry included
        bool have = false; A e = A(0);
        if (lane > 0) { e = cur.xe; have = true; if (have_c && !cur.fe) e = ScanOp<A, OP>::apply(c, cur.xe); }
        else if (have_c) { e = c; have = true; }
        s_prex[slot * 8 + lane] = scan_bits<A>(e);
        s_preh[slot * 8 + lane] = have ? 1u : 0u;
      }
      __syncwarp();
      if (lane == 0) sp_mbar_arrive(bar_pre + 8 * slot);
    }
    return;
  }

  // ---------------- data warps ----------------
  bool owe = false; int tp = 0;
  A xp[SG_ITEMS]; unsigned hmask_p = 0; bool fe_p = false; A xe_p = A(0);
  int4* my_out = s_out + wrp * 128;
  auto settle = [&](int itp) {                                           // finish tile tp (iteration itp): prefix, results, store
    const int slot = itp & 1;
    sp_mbar_wait(bar_pre + 8 * slot, (unsigned)(itp >> 1) & 1u);
    const bool have_e = s_preh[slot * 8 + wrp] != 0u; const A e = scan_unbits<A>(s_prex[slot * 8 + wrp]);
    bool have = false; A pre = A(0);
    if (lane > 0) { pre = xe_p; have = true; if (have_e && !fe_p) pre = ScanOp<A, OP>::apply(e, xe_p); }
    else if (have_e) { pre = e; have = true; }
    const int nopen = !have ? 0 : (hmask_p ? __ffs((int)hmask_p) - 1 : SG_ITEMS);
    unsigned long long ob[SG_ITEMS];
#pragma unroll
    for (int k = 0; k < SG_ITEMS; ++k) {
      const A r = k < nopen ? ScanOp<A, OP>::apply(pre, xp[k]) : xp[k];
      ob[k] = (std::is_same<A, double>::value ? (unsigned long long)__double_as_longlong((double)r) : (unsigned long long)(long long)r) ^ keyflip;
    }
    if (lane == 0) sp_store_read_done();                                 // this warp's previous store has read its slice
    __syncwarp();
    sp_write8(s_out, t, ob);
    sp_fence_async();
    __syncwarp();
    if (lane == 0) sp_tensor_store(&tm_out32, tp * SG_THREADS + wrp * 32, sp_saddr(my_out));
  };
  int it = 0;
  for (;; ++it) {
    const int tile = it * grid + cta;
    if (tile >= nfull) break;
    const int s = it % SP_STAGES, slot = it & 1;
    sp_mbar_wait(bar_full + 8 * s, (unsigned)(it / SP_STAGES) & 1u);
    unsigned long long gl[SG_ITEMS], raw[SG_ITEMS];
    const int4* lab_s = s_in + (2 * s) * (SG_TILE / 2);
    if (LB == 8) sp_read8(lab_s, t, gl);
    else {
      const int r1 = (t >> 2) & 1;
      const int4 u0 = lab_s[2 * t + r1], u1 = lab_s[2 * t + (1 ^ r1)];
      gl[0] = (unsigned long long)(long long)u0.x; gl[1] = (unsigned long long)(long long)u0.y; gl[2] = (unsigned long long)(long long)u0.z; gl[3] = (unsigned long long)(long long)u0.w;
      gl[4] = (unsigned long long)(long long)u1.x; gl[5] = (unsigned long long)(long long)u1.y; gl[6] = (unsigned long long)(long long)u1.z; gl[7] = (unsigned long long)(long long)u1.w;
    }
    sp_read8(s_in + (2 * s + 1) * (SG_TILE / 2), t, raw);
    const long long base = (long long)tile * SG_TILE + (long long)t * SG_ITEMS;
    long long prev = 0;
    const long long left = __shfl_up_sync(0xffffffffu, (long long)gl[SG_ITEMS - 1], 1);
    if (lane > 0) prev = left;
    else if (t > 0) {                                                    // the last label of the row before this thread's
      if (LB == 8) { const int4 q = lab_s[4 * (t - 1) + (3 ^ (((t - 1) >> 1) & 3))]; prev = (long long)(((unsigned long long)(unsigned)q.z) | ((unsigned long long)(unsigned)q.w << 32)); }
      else { const int4 q = lab_s[2 * (t - 1) + (1 ^ (((t - 1) >> 2) & 1))]; prev = (long long)q.w; }
    } else if (tile > 0) {
      const int4 q = s_left[s];
      prev = LB == 8 ? (long long)(((unsigned long long)(unsigned)q.z) | ((unsigned long long)(unsigned)q.w << 32)) : (long long)q.w;
    }
    A x[SG_ITEMS]; unsigned hmask = 0;
    A tx = A(0); bool uns = false, bad = false;
#pragma unroll
    for (int k = 0; k < SG_ITEMS; ++k) {
      const long long g = (long long)gl[k];
      const bool head = (k == 0 && base == 0) || g != prev;
      uns |= g < prev && !(k == 0 && base == 0);
      bad |= (unsigned long long)g >= (unsigned long long)p.size;
      prev = g;
      A v;
      if (VK == 2) { double d = __longlong_as_double((long long)raw[k]); if (NANSKIP && d != d) d = OP == SC_SUM ? 0.0 : 1.0; v = (A)d; }
      else v = (A)(long long)(raw[k] ^ keyflip);
      tx = (k == 0 || head) ? v : ScanOp<A, OP>::apply(tx, v);
      hmask |= head ? (1u << k) : 0u;
      x[k] = tx;
    }
    if (uns) *p.unsorted_w = 1;
    if (bad) p.status[AGC_ST_BADLABEL] = 1;
    bool fi = hmask != 0u; A xi = tx;                                    // inclusive over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const bool ft = __shfl_up_sync(0xffffffffu, (int)fi, o);
      const A xt = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) { if (!fi) xi = ScanOp<A, OP>::apply(xt, xi); fi = fi | ft; }
    }
    const bool fe = __shfl_up_sync(0xffffffffu, (int)fi, 1); const A xe = __shfl_up_sync(0xffffffffu, xi, 1);
    if (lane == 31) {                                                    // (every lane of the warp is past its reads of stage s)
      s_aggf[slot * 8 + wrp] = fi ? 1u : 0u; s_aggx[slot * 8 + wrp] = scan_bits<A>(xi);
      sp_mbar_arrive(bar_agg + 8 * slot);
    }
    if (owe) settle(it - 1);
#pragma unroll
    for (int k = 0; k < SG_ITEMS; ++k) xp[k] = x[k];
    hmask_p = hmask; fe_p = fe; xe_p = xe; tp = tile; owe = true;
  }
  if (owe) settle(it - 1);
  if (lane == 0) sp_store_done();
}

// single block: segmented scan over the tile aggregates -> carry for tile t = aggregate of tiles [0, t).
// (f, x) (+) (g, y) = (f | g, g ? y : op(x, y)).  Tile 0 always starts with a head flag, so every exclusive prefix
// with t > 0 is well defined.  The tiles are walked in coalesced chunks of 1024 with a running carry; the next
// chunk's loads are issued before the current chunk is scanned.
template <class A, int OP>
__global__ void __launch_bounds__(1024) k_seg_scan_tiles(const unsigned char* tile_f, const A* tile_x, int ntiles, unsigned char* carry_ok, A* carry_x,
                                                         const int* run_flag, int run_if) {
  if (run_flag && (*run_flag != 0) != (run_if != 0)) return;
  __shared__ unsigned char sf[32];
  __shared__ long long sx[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  bool rf = true; A rx = A(0);                               // running carry = aggregate of all earlier chunks
  bool nf = true; A nx = A(0);
  if (t < ntiles) { nf = tile_f[t]; nx = tile_x[t]; }
  for (int c0 = 0; c0 < ntiles; c0 += 1024) {
    const int idx = c0 + t;
    const bool f = nf; const A x = nx;
    const int nidx = idx + 1024;
    nf = true; nx = A(0);
    if (nidx < ntiles) { nf = tile_f[nidx]; nx = tile_x[nidx]; }
    // inclusive segmented scan inside the warp
    bool fi = f; A xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const bool ft = __shfl_up_sync(0xffffffffu, (int)fi, o);
      const A xt = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) { if (!fi) xi = ScanOp<A, OP>::apply(xt, xi); fi = fi | ft; }
    }
    if (lane == 31) { sf[w] = fi; ((A*)sx)[w] = xi; }
    bool fe = __shfl_up_sync(0xffffffffu, (int)fi, 1); A xe = __shfl_up_sync(0xffffffffu, xi, 1);   // exclusive in warp (lane > 0)
    __syncthreads();
    if (w == 0) {                                            // scan of the 32 warp aggregates
      bool af = sf[lane]; A ax = ((A*)sx)[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const bool ft = __shfl_up_sync(0xffffffffu, (int)af, o);
        const A xt = __shfl_up_sync(0xffffffffu, ax, o);
        if (lane >= o) { if (!af) ax = ScanOp<A, OP>::apply(xt, ax); af = af | ft; }
      }
      sf[lane] = af; ((A*)sx)[lane] = ax;                    // inclusive over warps
    }
    __syncthreads();
    // exclusive prefix of this thread inside the chunk: warps before (+) lanes before
    bool pf = false; A px = A(0); bool havep = false;
    if (w > 0) { pf = sf[w - 1]; px = ((A*)sx)[w - 1]; havep = true; }
    if (lane > 0) {
      if (havep && !fe) { px = ScanOp<A, OP>::apply(px, xe); pf = pf | fe; }
      else { px = xe; pf = fe | pf; }
      havep = true;
    }
    // with the running carry of earlier chunks
    bool ef; A ex;
    if (havep) { ef = pf | (c0 > 0 && rf); ex = (c0 > 0 && !pf) ? ScanOp<A, OP>::apply(rx, px) : px; }
    else { ef = rf; ex = rx; }
    (void)ef;
    if (idx < ntiles) { carry_ok[idx] = idx > 0; carry_x[idx] = ex; }
    // new running carry = old (+) chunk aggregate (inclusive value of the last warp)
    const bool cf = sf[31]; const A cx = ((A*)sx)[31];
    if (c0 == 0) { rf = cf; rx = cx; }
    else { if (!cf) rx = ScanOp<A, OP>::apply(rx, cx); else rx = cx; rf = rf | cf; }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// func='sort' helpers
// ---------------------------------------------------------------------------------------------
// range[0] / range[1] (may be NULL): smallest / largest key, for the range-compressed sort of integer values
__global__ void k_value_keys(const void* a, int a_dtype, long long n, int reverse, unsigned long long* keys, unsigned* pos,
                             unsigned long long* range) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long lo = 0xffffffffffffffffull, hi = 0ull;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned long long u;
    bool nan = false;
    // reverse: the reference sorts -a (aggregate_numpy.py:210-214, np.lexsort((-a, group_idx))), and for integers numpy's
    // negation WRAPS in the array's own width: the most negative value of a signed type and 0 of an unsigned type are their
    // own negatives and stay in front.  Floats: -a is exact, the key's order simply flips; NaNs sort last either way.
    if (a_dtype == AGC_DT_F32) { float v = ((const float*)a)[i]; nan = v != v; u = (unsigned long long)key_of((double)v) ^ 0x8000000000000000ull; if (reverse) u = ~u; }
    else if (a_dtype == AGC_DT_F64) { double v = ((const double*)a)[i]; nan = v != v; u = (unsigned long long)key_of(v) ^ 0x8000000000000000ull; if (reverse) u = ~u; }
    else if (a_dtype == AGC_DT_U64) { u = ((const unsigned long long*)a)[i]; if (reverse) u = 0ull - u; }
    else {
      long long v = load_int(a, a_dtype, i);
      if (reverse) {
        const int bits = 8 * dtype_bytes(a_dtype);
        const bool uns = a_dtype == AGC_DT_U8 || a_dtype == AGC_DT_U16 || a_dtype == AGC_DT_U32 || a_dtype == AGC_DT_BOOL;
        unsigned long long w = 0ull - (unsigned long long)v;
        if (bits < 64) {
          w &= (1ull << bits) - 1ull;
          if (!uns && (w >> (bits - 1))) w |= ~((1ull << bits) - 1ull);     // sign-extend from the array's width
        }
        v = (long long)w;
      }
      u = (unsigned long long)v ^ 0x8000000000000000ull;
    }
    if (nan) u = 0xffffffffffffffffull;                  // NaNs sort last in both directions (np.lexsort)
    keys[i] = u; pos[i] = (unsigned)i;
    lo = u < lo ? u : lo; hi = u > hi ? u : hi;
  }
  if (range) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
      lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(range, lo); atomicMax(range + 1, hi); }
  }
}
// Range-compressed value sort: integer values rarely use all 64 bits of their key (the C4 workload: int64 holding 32-bit
// numbers), and every 8 bits the span max - min does not reach are a radix pass (6.5 ms at N = 2^28) that moves nothing.  The
// passes sort (key - min); pass p runs iff 8p < bits(max - min).  The skipped passes are a SUFFIX, so which buffer holds the
// result follows from the number of passes run: flags[8] says whether that number is odd (then k_copy_if brings the
// positions to the buffer the host expects).
__global__ void k_range_init(unsigned long long* range) { range[0] = 0xffffffffffffffffull; range[1] = 0ull; }
__global__ void k_pass_flags(const unsigned long long* range, int* flags) {
  const unsigned long long span = range[1] >= range[0] ? range[1] - range[0] : 0ull;
  const int nbits = span ? 64 - __clzll((long long)span) : 0;
  const int npass = (nbits + 7) / 8;
  if (threadIdx.x < 8) flags[threadIdx.x] = (int)threadIdx.x < npass;
  if (threadIdx.x == 8) flags[8] = npass & 1;
}
__global__ void k_copy_if(const unsigned* src, unsigned* dst, long long n, const int* flag) {
  if (*flag == 0) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
__global__ void k_keys_of_positions(Indexer ix, long long n, const unsigned* pos, unsigned* keys) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    long long g = ix.group(pos[i]);
    keys[i] = g < 0 ? 0u : (unsigned)g;
  }
}
__global__ void k_sort_place(const void* a, int eb, long long n, const unsigned* slot_pos, const unsigned* src_pos, void* out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned d = slot_pos[i], s = src_pos[i];
    switch (eb) {
      case 1: ((unsigned char*)out)[d] = ((const unsigned char*)a)[s]; break;
      case 2: ((unsigned short*)out)[d] = ((const unsigned short*)a)[s]; break;
      case 4: ((unsigned*)out)[d] = ((const unsigned*)a)[s]; break;
      default: ((unsigned long long*)out)[d] = ((const unsigned long long*)a)[s]; break;
    }
  }
}
__global__ void k_gather_by_pos(const void* a, int eb, long long n, const unsigned* src_pos, void* staged) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned s = src_pos[i];
    if (eb == 8) ((unsigned long long*)staged)[i] = gather64(a, s);
    else ((unsigned*)staged)[i] = gather32(a, s);
  }
}
__global__ void k_order_widen(const unsigned* p0, const unsigned* p1, const int* unsorted, long long n, long long* order) {
  const unsigned* pos = *unsorted ? p1 : p0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) order[i] = pos[i];
}
__global__ void k_csr_offsets(const unsigned* keys0, const unsigned* keys1, const int* unsorted, long long n, long long size,
                              long long* offsets) {
  const unsigned* keys = *unsorted ? keys1 : keys0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g <= size; g += stride) {
    long long lo = 0, hi = n;                               // lower_bound(keys, g)
    while (lo < hi) { long long mid = (lo + hi) >> 1; if ((long long)keys[mid] < g) lo = mid + 1; else hi = mid; }
    offsets[g] = lo;
  }
}
__global__ void k_pos_select(const unsigned* p0, const unsigned* p1, const int* unsorted, long long n, unsigned* out) {
  const unsigned* pos = *unsorted ? p1 : p0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) out[i] = pos[i];
}

// ---------------------------------------------------------------------------------------------
// workspace carving
// ---------------------------------------------------------------------------------------------
inline int64_t al256(int64_t x) { return (x + 255) / 256 * 256; }
struct ScanWs {
  int64_t k32[2], v[2], k64[2], v2[2], slot, gath, hist, tsum, tile_f, tile_x, carry_f, carry_x, desc, flag, stage, total;
};
inline ScanWs scan_ws_layout(long long n, bool want_sort_op) {
  ScanWs w; int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o += al256(bytes); return r; };
  const int nb = rs_blocks(n > 0 ? n : 1);
  w.k32[0] = take(n * 4); w.k32[1] = take(n * 4); w.v[0] = take(n * 4); w.v[1] = take(n * 4);
  if (want_sort_op) { w.k64[0] = take(n * 8); w.k64[1] = take(n * 8); w.v2[0] = take(n * 4); w.v2[1] = take(n * 4); w.slot = take(n * 4); }
  else { w.k64[0] = w.k64[1] = w.v2[0] = w.v2[1] = w.slot = 0; }
  w.gath = want_sort_op ? 0 : take(n * 8);
  w.stage = want_sort_op ? 0 : take(n * 8);
  w.hist = take(256ll * nb * 4); w.tsum = take((xs_tiles(256ll * nb) + 1) * 4);
  const long long ntiles = (n + SG_TILE - 1) / SG_TILE + 1;
  w.tile_f = take(ntiles); w.tile_x = take(ntiles * 8); w.carry_f = take(ntiles); w.carry_x = take(ntiles * 8);
  w.desc = take(ntiles * 16);
  w.flag = take(256);
  w.total = o;
  return w;
}
inline int64_t scan_workspace_bytes(int op, int64_t n, int64_t, int) { return scan_ws_layout(n, op == AGC_OP_SORT).total; }
inline int64_t sort_workspace_bytes(int64_t n, int64_t) { return scan_ws_layout(n, false).total; }

inline int bits_for(long long size) { int b = 1; while (b < 32 && (1ll << b) < size) ++b; return (b + 7) / 8 * 8; }

inline int scan_grid(long long n, int sms) { long long b = (n + 255) / 256; long long cap = (long long)sms * 16; return (int)(b < 1 ? 1 : (b > cap ? cap : b)); }

// generate group keys, detect sortedness, sort (key, pos) by key unless sorted.  Leaves results addressed
// through (keys[0|1], pos[0|1]) selected by *flag on the device.
inline int group_sorted_pairs(const Indexer& ix, long long n, long long size, char* ws, const ScanWs& L, int* status,
                              bool assume_sorted, cudaStream_t st, int sms, const unsigned* (&keys)[2], const unsigned* (&pos)[2],
                              int*& flag, bool flag_ready = false) {
  if (n >= (1ll << 32) || size > (1ll << 32)) { g_scan_err = "N->N / sort paths support n < 2^32 and size <= 2^32"; return -21; }
  flag = (int*)(ws + L.flag);
  // flag_ready: the direct scan already set *flag (1 = labels not sorted); everything below then only runs when it is 1
  if (!flag_ready) AGC_COUNT_LAUNCH(1), k_zero_int<<<1, 1, 0, st>>>(flag);
  unsigned* k0 = (unsigned*)(ws + L.k32[0]); unsigned* p0 = (unsigned*)(ws + L.v[0]);
  if (n > 0) AGC_COUNT_LAUNCH(1), k_group_keys<<<scan_grid(n, sms), 256, 0, st>>>(ix, n, k0, p0, status, flag, flag_ready ? 1 : 0);
  SortBufs b; b.k[0] = k0; b.k[1] = ws + L.k32[1]; b.v[0] = p0; b.v[1] = (unsigned*)(ws + L.v[1]);
  b.hist = (unsigned*)(ws + L.hist); b.tile_sums = (unsigned*)(ws + L.tsum);
  int res = 0;
  if (!assume_sorted) res = radix_sort_pairs<unsigned>(b, 0, n, 0, bits_for(size), flag, st);
  else AGC_COUNT_LAUNCH(1), k_report_unsorted<<<1, 1, 0, st>>>(flag, status);
  keys[0] = k0; pos[0] = p0; keys[1] = (const unsigned*)b.k[res]; pos[1] = b.v[res];
  return 0;
}

template <class A, int OP>
inline void run_seg_scan(const ScanParams& p, char* ws, const ScanWs& L, cudaStream_t st) {
  if (p.n == 0) return;
  const int ntiles = (int)((p.n + SG_TILE - 1) / SG_TILE);
  unsigned char* tf = (unsigned char*)(ws + L.tile_f); A* tx = (A*)(ws + L.tile_x);
  unsigned char* cf = (unsigned char*)(ws + L.carry_f); A* cx = (A*)(ws + L.carry_x);
  AGC_COUNT_LAUNCH(1), k_seg_scan<A, OP, 0><<<ntiles, SG_THREADS, 0, st>>>(p, tf, tx, nullptr, nullptr);
  AGC_COUNT_LAUNCH(1), k_seg_scan_tiles<A, OP><<<1, 1024, 0, st>>>(tf, tx, ntiles, cf, cx, p.only_if_unsorted ? p.unsorted : nullptr, 1);
  AGC_COUNT_LAUNCH(1), k_seg_scan<A, OP, 1><<<ntiles, SG_THREADS, 0, st>>>(p, nullptr, nullptr, cf, cx);
  if (p.stage) {                                             // (every kernel below is a no-op unless the labels really were unsorted)
    const int eb = dtype_bytes(p.out_dtype);
    unsigned* hist = (unsigned*)(ws + L.hist); unsigned* tsum = (unsigned*)(ws + L.tsum);
    // free by now: the pair of sort buffers that does not hold the result, and the gathered inputs
    unsigned* pos_part = (unsigned*)(ws + (p.pos[1] == (const unsigned*)(ws + L.v[1]) ? L.k32[0] : L.k32[1]));
    if (eb == 8) unsort_windowed<unsigned long long>(p.pos[1], (const unsigned long long*)p.stage, p.n, (unsigned long long*)p.out, pos_part, (unsigned long long*)p.gath, hist, tsum, p.unsorted, st);
    else unsort_windowed<unsigned>(p.pos[1], (const unsigned*)p.stage, p.n, (unsigned*)p.out, pos_part, (unsigned*)p.gath, hist, tsum, p.unsorted, st);
  }
}

// AGC_SCAN_CHAIN=0 (or agc_tuning key 6 = 0) selects the three-kernel direct scan instead of the chained single pass
static int g_scan_chain = -1;
inline bool scan_chain_enabled() {
  if (g_scan_chain < 0) { const char* e = getenv("AGC_SCAN_CHAIN"); g_scan_chain = (e && atoi(e) == 0) ? 0 : 1; }
  return g_scan_chain != 0;
}
// tensor map of an array of 8-byte elements seen as rows of 8 (64 bytes); one box = one tile = 256 rows
typedef CUresult (*sp_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline sp_encode_fn sp_encoder() {
  static sp_encode_fn fn = [] {
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (sp_encode_fn)f;
  }();
  return fn;
}
inline bool sp_tile_map(CUtensorMap* tm, const void* base, int nfull, int box_rows = SG_THREADS, int elem_bytes = 8) {
  sp_encode_fn enc = sp_encoder();
  if (!enc || nfull <= 0) return false;
  // rows of 8 elements: 64 bytes with the 64-byte swizzle for 8-byte elements, 32 bytes with the 32-byte swizzle for int32 labels
  const cuuint64_t dims[2] = {8, (cuuint64_t)nfull * SG_THREADS};
  const cuuint64_t strides[1] = {(cuuint64_t)(8 * elem_bytes)};
  const cuuint32_t box[2] = {8, (cuuint32_t)box_rows}, estr[2] = {1, 1};
  return enc(tm, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, elem_bytes == 8 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <class A, int OP, class LT>
inline void run_seg_scan_direct(const ScanParams& p, char* ws, const ScanWs& L, cudaStream_t st) {
  if (p.n == 0) return;
  const int ntiles = (int)((p.n + SG_TILE - 1) / SG_TILE);
  unsigned char* tf = (unsigned char*)(ws + L.tile_f); A* tx = (A*)(ws + L.tile_x);
  unsigned char* cf = (unsigned char*)(ws + L.carry_f); A* cx = (A*)(ws + L.carry_x);
  // (deterministic mode keeps the three-kernel variant: its order of floating additions does not depend on which tiles
  //  had published a prefix when a look-back came by)
  if (scan_chain_enabled() && !(p.flags & AGC_F_DETERMINISTIC)) {          // one pass, 24 B/element: tile descriptors + warp-parallel look-back
    cudaMemsetAsync(ws + L.desc, 0, (size_t)ntiles * 16, st);
    int nfull = 0;
    // full tiles of the all-64-bit case (the C4 configuration): the lean kernel, everything decided at compile time
    {
      const bool i64v = p.a_dtype == AGC_DT_I64, u64v = p.a_dtype == AGC_DT_U64, f64v = p.a_dtype == AGC_DT_F64;
      const bool ints = std::is_same<A, long long>::value && (i64v || u64v) && (p.out_dtype == AGC_DT_I64 || p.out_dtype == AGC_DT_U64);
      const bool flts = std::is_same<A, double>::value && f64v && p.out_dtype == AGC_DT_F64 && (OP == SC_SUM || OP == SC_PROD);
      static const bool lean_on = !(getenv("AGC_SCAN_LEAN") && atoi(getenv("AGC_SCAN_LEAN")) == 0);
      constexpr int LBYTES = (int)sizeof(LT);                // int64 or int32 labels; the one-tile-per-CTA fallback takes int64 only
      if (lean_on && (ints || flts) && !((uintptr_t)p.a & 15) && !((uintptr_t)p.out & 15)) {
        nfull = (int)(p.n / SG_TILE);
        if (nfull > 0) {
          unsigned long long* d = (unsigned long long*)(ws + L.desc);
          const bool ns = (p.flags & AGC_F_NANSKIP) != 0;
          AGC_COUNT_LAUNCH(1);
          // AGC_SCAN_WS=0: the one-tile-per-CTA kernel (k_seg_scan_lean) instead of the persistent warp-specialised one
          static const int occ = getenv("AGC_SCAN_OCC") ? atoi(getenv("AGC_SCAN_OCC")) : 6;
          static const bool ws_on = !(getenv("AGC_SCAN_WS") && atoi(getenv("AGC_SCAN_WS")) == 0);
          CUtensorMap tml, tmv, tmo32;                           // tiles of 256 rows x 64 bytes (64-byte swizzle); tmo32: one warp's 32 rows of the result
          const bool ws_ok = ws_on && sp_tile_map(&tml, p.labels, nfull, SG_THREADS, LBYTES) && sp_tile_map(&tmv, p.a, nfull) && sp_tile_map(&tmo32, p.out, nfull, 32);
          // cooperative launch: the runtime refuses a grid that cannot be resident as a whole (then the one-tile-per-CTA kernel
          // runs) and never starts part of it -- the round-robin tile order of the kernel depends on exactly that
#define AGC_PIPE(VK_, NS_)                                                                                              \
          { static std::atomic<unsigned long long> attr_done{0};                                                           \
            int per_sm = 0;                                                                                                \
            if (ensure_dyn_smem(k_seg_scan_ws<A, OP, VK_, NS_, LBYTES>, WS_SMEM, attr_done))                                       \
              cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_seg_scan_ws<A, OP, VK_, NS_, LBYTES>, WS_THREADS, WS_SMEM); \
            const long long cap = (long long)per_sm * device_sms();                                                        \
            const int pgrid = (int)(nfull < cap ? nfull : cap);                                                            \
            launched = false;                                                                                              \
            if (pgrid > 0) {                                                                                               \
              ScanParams pa = p; unsigned long long* da = d; int nf = nfull;                                               \
              void* args[] = {&pa, &da, &nf, &tml, &tmv, &tmo32};                                                          \
              launched = cudaLaunchCooperativeKernel((const void*)k_seg_scan_ws<A, OP, VK_, NS_, LBYTES>, dim3(pgrid), dim3(WS_THREADS), args, WS_SMEM, st) == cudaSuccess; \
              if (!launched) cudaGetLastError();                                                                           \
            } }
#define AGC_LEAN(VK_, NS_)                                                                                              \
          { bool launched = ws_ok;                                                                                         \
            if (ws_ok) AGC_PIPE(VK_, NS_)                                                                                  \
            if (launched) {}                                                                                               \
            else if (LBYTES != 8) nfull = 0;                       /* int32 labels: the general chained kernel takes every tile */ \
            else if (occ == 5) k_seg_scan_lean<A, OP, VK_, NS_, 5><<<nfull, SG_THREADS, 0, st>>>(p, d);                     \
            else k_seg_scan_lean<A, OP, VK_, NS_, 6><<<nfull, SG_THREADS, 0, st>>>(p, d); }
          if (flts) { if (ns) AGC_LEAN(2, true) else AGC_LEAN(2, false) }
          else if (u64v) AGC_LEAN(1, false)
          else AGC_LEAN(0, false)
#undef AGC_LEAN
#undef AGC_PIPE
        }
      }
    }
    if (nfull < ntiles) {
      ScanParams q = p; q.tile0 = nfull;
      AGC_COUNT_LAUNCH(1), k_seg_scan_direct<A, OP, LT, 2><<<ntiles - nfull, SG_THREADS, 0, st>>>(q, nullptr, (A*)(ws + L.desc), nullptr, nullptr);
    }
    return;
  }
  AGC_COUNT_LAUNCH(1), k_seg_scan_direct<A, OP, LT, 0><<<ntiles, SG_THREADS, 0, st>>>(p, tf, tx, nullptr, nullptr);
  AGC_COUNT_LAUNCH(1), k_seg_scan_tiles<A, OP><<<1, 1024, 0, st>>>(tf, tx, ntiles, cf, cx, p.unsorted, 0);
  AGC_COUNT_LAUNCH(1), k_seg_scan_direct<A, OP, LT, 1><<<ntiles, SG_THREADS, 0, st>>>(p, nullptr, nullptr, cf, cx);
}
template <class A, int OP>
inline void run_seg_scan_any(const ScanParams& p, char* ws, const ScanWs& L, cudaStream_t st, int direct_lt) {
  if (direct_lt == 8) run_seg_scan_direct<A, OP, long long>(p, ws, L, st);
  else if (direct_lt == 4) run_seg_scan_direct<A, OP, int>(p, ws, L, st);
  else run_seg_scan<A, OP>(p, ws, L, st);
}
inline void dispatch_seg_scan(const agc_request_t* r, const ScanParams& p, char* ws, const ScanWs& L, cudaStream_t st, int direct_lt) {
  const bool facc = acc_is_float(r->op, r->a_dtype, r->out_dtype);
  switch (r->op) {
    case AGC_OP_CUMSUM: if (facc) run_seg_scan_any<double, SC_SUM>(p, ws, L, st, direct_lt); else run_seg_scan_any<long long, SC_SUM>(p, ws, L, st, direct_lt); break;
    case AGC_OP_CUMPROD: if (facc) run_seg_scan_any<double, SC_PROD>(p, ws, L, st, direct_lt); else run_seg_scan_any<long long, SC_PROD>(p, ws, L, st, direct_lt); break;
    case AGC_OP_CUMMAX: run_seg_scan_any<long long, SC_MAX>(p, ws, L, st, direct_lt); break;
    default: run_seg_scan_any<long long, SC_MIN>(p, ws, L, st, direct_lt); break;
  }
}

inline int run_scan_op(const agc_request_t* r, cudaStream_t st, int sms) {
  const long long n = r->n;
  if (r->a == nullptr && n > 0) { g_scan_err = "N->N operations need an array a"; return -22; }
  const bool sort_op = r->op == AGC_OP_SORT;
  const ScanWs L = scan_ws_layout(n, sort_op);
  if (r->workspace_bytes < L.total) { g_scan_err = "workspace too small"; return -8; }
  char* ws = (char*)r->workspace;
  Indexer ix; ix.ix = r->index; ix.size = r->size; ix.n = n;
  const unsigned* keys[2]; const unsigned* pos[2]; int* flag = nullptr;
  if (n >= (1ll << 32) || r->size > (1ll << 32)) { g_scan_err = "N->N / sort paths support n < 2^32 and size <= 2^32"; return -21; }

  ScanParams p;
  memset(&p, 0, sizeof(p));
  p.a = r->a; p.a_dtype = r->a_dtype; p.out_dtype = r->out_dtype; p.out = r->out; p.n = n; p.flags = r->flags;
  p.fill_f = r->fill_f64; p.fill_i = r->fill_i64;
  p.a_float = dtype_is_float(r->a_dtype); p.a_u64 = r->a_dtype == AGC_DT_U64;
  p.labels = r->index.labels; p.size = r->size; p.status = r->status;
  { static const int lb_sleep = (getenv("AGC_LB_SLEEP") ? atoi(getenv("AGC_LB_SLEEP")) & 0xffff : 0) | ((getenv("AGC_DESC_GPU") && atoi(getenv("AGC_DESC_GPU"))) ? (1 << 30) : 0); p.lb_sleep = lb_sleep; }

  // ---- flat int32/int64 labels: try the direct (already sorted) scan first -------------------------------
  int direct_lt = 0;
  if (!sort_op && n > 0 && r->index.form == AGC_FORM_FLAT && !((uintptr_t)r->index.labels & 15)) {
    if (r->index.dtype == AGC_DT_I64) direct_lt = 8; else if (r->index.dtype == AGC_DT_I32) direct_lt = 4;
  }
  if (direct_lt) {
    flag = (int*)(ws + L.flag);
    AGC_COUNT_LAUNCH(1), k_zero_int<<<1, 1, 0, st>>>(flag);
    if (n >= (1ll << 20)) {
      if (direct_lt == 8) AGC_COUNT_LAUNCH(1), k_sorted_sample<long long><<<16, 256, 0, st>>>((const long long*)r->index.labels, n, flag);
      else AGC_COUNT_LAUNCH(1), k_sorted_sample<int><<<16, 256, 0, st>>>((const int*)r->index.labels, n, flag);
    }
    p.unsorted = flag; p.unsorted_w = flag;
    dispatch_seg_scan(r, p, ws, L, st, direct_lt);
    if (r->flags & AGC_F_ASSUME_SORTED) { AGC_COUNT_LAUNCH(1), k_report_unsorted<<<1, 1, 0, st>>>(flag, r->status); return 0; }
  }

  int rc = group_sorted_pairs(ix, n, r->size, ws, L, r->status, (r->flags & AGC_F_ASSUME_SORTED) != 0, st, sms, keys, pos, flag, direct_lt != 0);
  if (rc) return rc;
  if (n == 0) return 0;
  if (sort_op) {
    // slots: positions in group order
    unsigned* slot = (unsigned*)(ws + L.slot);
    AGC_COUNT_LAUNCH(1), k_pos_select<<<scan_grid(n, sms), 256, 0, st>>>(pos[0], pos[1], flag, n, slot);
    // elements ordered by (group, value): sort by value key, then stably by group
    SortBufs b; b.k[0] = ws + L.k64[0]; b.k[1] = ws + L.k64[1]; b.v[0] = (unsigned*)(ws + L.v2[0]); b.v[1] = (unsigned*)(ws + L.v2[1]);
    b.hist = (unsigned*)(ws + L.hist); b.tile_sums = (unsigned*)(ws + L.tsum);
    // integer values: sort (key - min) and skip the passes above bits(max - min) on the device (see k_pass_flags)
    static const bool range_on = !(getenv("AGC_SORT_RANGE") && atoi(getenv("AGC_SORT_RANGE")) == 0);
    const bool ranged = range_on && !dtype_is_float(r->a_dtype);
    unsigned long long* range = (unsigned long long*)(ws + L.flag + 64);
    int* pass_flags = (int*)(ws + L.flag + 128);
    if (ranged) AGC_COUNT_LAUNCH(1), k_range_init<<<1, 1, 0, st>>>(range);
    AGC_COUNT_LAUNCH(1), k_value_keys<<<scan_grid(n, sms), 256, 0, st>>>(r->a, r->a_dtype, n, (r->flags & AGC_F_REVERSE) ? 1 : 0,
                                                    (unsigned long long*)b.k[0], b.v[0], ranged ? range : nullptr);
    if (ranged) AGC_COUNT_LAUNCH(1), k_pass_flags<<<1, 32, 0, st>>>(range, pass_flags);
    int lo_bit = 0;
    if (r->a_dtype == AGC_DT_F32) lo_bit = 24;              // f32 widened to f64: low 29 mantissa bits are zero
    int res = radix_sort_pairs<unsigned long long>(b, 0, n, lo_bit, 64, nullptr, st, ranged ? pass_flags : nullptr, ranged ? range : nullptr);
    if (ranged) AGC_COUNT_LAUNCH(1), k_copy_if<<<scan_grid(n, sms), 256, 0, st>>>(b.v[res ^ 1], b.v[res], n, pass_flags + 8);
    // group keys of the value-ordered positions, then a stable sort by group
    SortBufs g; g.k[0] = ws + L.k32[0]; g.k[1] = ws + L.k32[1]; g.v[0] = b.v[res]; g.v[1] = b.v[res ^ 1];
    g.hist = b.hist; g.tile_sums = b.tile_sums;
    AGC_COUNT_LAUNCH(1), k_keys_of_positions<<<scan_grid(n, sms), 256, 0, st>>>(ix, n, g.v[0], (unsigned*)g.k[0]);
    int res2 = radix_sort_pairs<unsigned>(g, 0, n, 0, bits_for(r->size), nullptr, st);
    // placement out[slot[i]] = a[src[i]]: for large inputs the gather goes to a staging buffer in sorted order (the value-key
    // buffers are free by now) and the window un-sort brings the values home, instead of a random gather AND a random scatter
    // per element (20 ms of the 62 at N = 2^28)
    const int ebs = dtype_bytes(r->a_dtype);
    static const bool win = !(getenv("AGC_UNSORT_WINDOW") && atoi(getenv("AGC_UNSORT_WINDOW")) == 0);
    if (win && n >= (1ll << 22) && (ebs == 8 || ebs == 4)) {
      void* stage = ws + L.k64[0]; void* val_part = ws + L.k64[1];
      unsigned* pos_part = (unsigned*)g.k[res2 ^ 1];           // the key buffer of the group sort that does not hold its result
      AGC_COUNT_LAUNCH(1), k_gather_by_pos<<<scan_grid(n, sms), 256, 0, st>>>(r->a, ebs, n, g.v[res2], stage);
      if (ebs == 8) unsort_windowed<unsigned long long>(slot, (const unsigned long long*)stage, n, (unsigned long long*)r->out, pos_part, (unsigned long long*)val_part, g.hist, g.tile_sums, nullptr, st);
      else unsort_windowed<unsigned>(slot, (const unsigned*)stage, n, (unsigned*)r->out, pos_part, (unsigned*)val_part, g.hist, g.tile_sums, nullptr, st);
      return 0;
    }
    AGC_COUNT_LAUNCH(1), k_sort_place<<<scan_grid(n, sms), 256, 0, st>>>(r->a, ebs, n, slot, g.v[res2], r->out);
    return 0;
  }
  p.keys[0] = keys[0]; p.keys[1] = keys[1]; p.pos[0] = pos[0]; p.pos[1] = pos[1]; p.unsorted = flag;
  p.only_if_unsorted = direct_lt != 0;
  p.gath = ws + L.gath;
  { static const bool win = !(getenv("AGC_UNSORT_WINDOW") && atoi(getenv("AGC_UNSORT_WINDOW")) == 0);
    const int eb = dtype_bytes(r->out_dtype);
    p.stage = (win && n >= (1ll << 22) && (eb == 8 || eb == 4)) ? (void*)(ws + L.stage) : nullptr; }
  if (r->op < AGC_OP_CUMSUM || r->op > AGC_OP_CUMMIN) { g_scan_err = "unknown N->N op"; return -2; }
  dispatch_seg_scan(r, p, ws, L, st, 0);
  return 0;
}

inline int group_order(const agc_index_t& index, int64_t n, int64_t size, long long* order, long long* offsets, void* workspace,
                       int* status, cudaStream_t st, int sms) {
  const ScanWs L = scan_ws_layout(n, false);
  char* ws = (char*)workspace;
  Indexer ix; ix.ix = index; ix.size = size; ix.n = n;
  const unsigned* keys[2]; const unsigned* pos[2]; int* flag = nullptr;
  int rc = group_sorted_pairs(ix, n, size, ws, L, status, false, st, sms, keys, pos, flag);
  if (rc) return rc;
  if (n > 0) AGC_COUNT_LAUNCH(1), k_order_widen<<<scan_grid(n, sms), 256, 0, st>>>(pos[0], pos[1], flag, n, order);
  if (offsets) AGC_COUNT_LAUNCH(1), k_csr_offsets<<<scan_grid(size + 1, sms), 256, 0, st>>>(keys[0], keys[1], flag, n, size, offsets);
  return 0;
}

}  // namespace agc
