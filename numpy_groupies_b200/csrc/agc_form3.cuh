// agc_form3.cuh -- Form 3 (`axis=` on the LAST axis of a C-contiguous f32 array, utils.py:411-432,495-498):
// every row of `a` is reduced with the same label vector, out[row][g].
//
// The label vector is a.shape[axis] long and is never expanded.  A work item is (block of RB rows) x (chunk of
// columns); a CTA keeps the RB x G accumulators of its item in shared memory (the integer formats of
// agc_fast.cuh: 64-bit fixed point sums, u32 counts, order-preserving s32 keys -- native ATOMS only), loads the
// labels of a column vector ONCE and streams the RB rows under them, so label traffic is 8/RB bytes per element
// and comes from L2.  Items are handed out through a device counter; each item ends with one RED per touched
// (row, group) into the global tables of agc_generic.cuh, which keeps Forms 3/4 on the same finalise path.
//
// Bound (round 2): sums run at ~575 Gelem/s = 2 elements / clk / SM whatever the fixed-point layout -- a 31-bit grid whose
// high word only moves on a carry (1.25 ATOMS per element, FMUL + F2I instead of f64 conversions) measured the SAME 1.87-1.95
// ms for 4096 x 2^18 as the two-ATOMS 40-bit grid kept here: what counts is the ONE shared-memory atomic per element that
// must RETURN the old word for the carry, and those retire at ~2 lanes / clk / SM (max / min, whose atomics return
// nothing, run at 3.8).  The same rate is what Form 1 sums hide behind their 12 B / element of HBM traffic.
#pragma once
#include "agc_fast.cuh"

namespace agc {

struct Form3Params {
  const void* labels; const float* a;
  long long rows, cols, G;
  Tables t;
  int* status;
  int flags, square, want_b1, want_cnt;
  long long chunk_vec;        // column vectors (4 columns) per item
  int n_col_chunks, n_items;
  int* counter;               // work-item counter (zeroed by the caller: status[7])
};

template <int MODE, class LT, int RB>
__global__ void __launch_bounds__(FAST_THREADS, 2) k_form3(const Form3Params p) {
  extern __shared__ unsigned sw[];
  __shared__ unsigned s_maxbits;
  __shared__ int s_item;
  constexpr int WORDS = MODE == FM_SUM ? 2 : (MODE == FM_SUMCNT ? 3 : 1);
  constexpr bool IS_SUM = MODE == FM_SUM || MODE == FM_SUMCNT;
  const int G = (int)p.G, T = RB * G;
  unsigned* w0 = sw; unsigned* w1 = sw + T; unsigned* w2 = sw + 2 * T;
  const unsigned init0 = MODE == FM_MAX ? 0x80000000u : (MODE == FM_MIN ? 0x7fffffffu : 0u);
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  const long long cvec = p.cols >> 2;
  bool bad = false;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) { s_item = atomicAdd(p.counter, 1); s_maxbits = 0u; }
    for (int i = threadIdx.x; i < WORDS * T; i += blockDim.x) sw[i] = (i < T) ? init0 : 0u;
    __syncthreads();
    const int item = s_item;
    if (item >= p.n_items) break;
    const long long row0 = (long long)(item / p.n_col_chunks) * RB;
    const long long v0 = (long long)(item % p.n_col_chunks) * p.chunk_vec;
    const long long v1 = v0 + p.chunk_vec < cvec ? v0 + p.chunk_vec : cvec;
    const int nr = p.rows - row0 < RB ? (int)(p.rows - row0) : RB;

    // reference exponent of this item from its first tile (see agc_fast.cuh)
    unsigned lo_bits = 0, span_bits = 0; double scale = 1.0, inv_scale = 1.0;
    if (IS_SUM) {
      unsigned m = 0;
      const long long v = v0 + threadIdx.x;
      if (v < v1) {
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          if (r >= nr) break;
          int4 av = ldg_stream16((const int4*)p.a + (row0 + r) * cvec + v);
          unsigned b[4] = {(unsigned)av.x & 0x7fffffffu, (unsigned)av.y & 0x7fffffffu, (unsigned)av.z & 0x7fffffffu, (unsigned)av.w & 0x7fffffffu};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            unsigned bb = b[k];
            if (p.square) { float f = __uint_as_float(bb); bb = __float_as_uint(f * f); }
            if (bb < 0x7f800000u && bb > m) m = bb;
          }
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

    auto update = [&](int gi, long long gflat, float x) {
      if (MODE == FM_LEN) {
        if (nanskip && x != x) return;
        atomicAdd(w0 + gi, 1u);
      } else if (MODE == FM_MAX || MODE == FM_MIN) {
        int k;
        if (x != x) { if (nanskip) return; k = MODE == FM_MAX ? 0x7fffffff : (int)0x80000000; }   // NaN wins
        else k = key32_of(x);
        if (MODE == FM_MAX) atomicMax((int*)w0 + gi, k); else atomicMin((int*)w0 + gi, k);
      } else {
        if (nanskip && x != x) return;
        if (p.square) x = x * x;
        unsigned ab = __float_as_uint(x) & 0x7fffffffu;
        if (ab - lo_bits < span_bits) fx_add(w0 + gi, w1 + gi, __double2ll_rn((double)x * scale));
        else if (ab != 0u) atomicAdd(p.t.t0f + gflat, (double)x);      // out of window: exact f64 RED
        if (MODE == FM_SUMCNT) atomicAdd(w2 + gi, 1u);
      }
    };

    for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
      long long g[4];
      LabelVec<LT>::load(p.labels, v, g);
      int4 av[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        av[r] = make_int4(0, 0, 0, 0);
        if (r < nr) av[r] = ldg_stream16((const int4*)p.a + (row0 + r) * cvec + v);
      }
      bool ok[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { ok[k] = g[k] >= 0 && g[k] < p.G; bad |= !ok[k]; }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (r >= nr) break;
        const float x[4] = {__int_as_float(av[r].x), __int_as_float(av[r].y), __int_as_float(av[r].z), __int_as_float(av[r].w)};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (ok[k]) update(r * G + (int)g[k], (row0 + r) * p.G + g[k], x[k]);
      }
    }
    __syncthreads();
    // merge the item's table into the global accumulators
    for (int i = threadIdx.x; i < nr * G; i += blockDim.x) {
      const long long gi = row0 * p.G + i;                     // (row0 + r) * G + g with i = r * G + g
      if (MODE == FM_LEN) {
        unsigned c = w0[i];
        if (c) atomicAdd((unsigned long long*)(p.t.t1 + gi), (unsigned long long)c);
      } else if (MODE == FM_MAX || MODE == FM_MIN) {
        int k = (int)w0[i];
        if (k != (int)init0) {
          long long k64;
          if (MODE == FM_MAX) k64 = k == 0x7fffffff ? AGC_EMPTY_MIN : key_of((double)unkey_f32(k));
          else k64 = k == (int)0x80000000 ? AGC_EMPTY_MAX : key_of((double)unkey_f32(k));
          if (MODE == FM_MAX) atomicMax(p.t.t0i + gi, k64); else atomicMin(p.t.t0i + gi, k64);
        }
      } else {
        long long q = (long long)(((unsigned long long)w1[i] << 32) | w0[i]);
        if (q) atomicAdd(p.t.t0f + gi, (double)q * inv_scale);
        if (MODE == FM_SUMCNT) {
          unsigned c = w2[i];
          if (c) {
            if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + gi), (unsigned long long)c);
            if (p.want_b1) p.t.b1[gi] = 1;
          }
        }
      }
    }
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.status[AGC_ST_REGIME] = 40 + MODE;
}

template <int MODE, class LT, int RB>
inline int launch_form3_rb(const Form3Params& p, size_t smem, cudaStream_t st, int sms) {
  auto kern = k_form3<MODE, LT, RB>;
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dyn_smem(kern, 110 * 1024, attr_done)) return -30;
  int grid = sms * 2;
  if (grid > p.n_items) grid = p.n_items;
  AGC_COUNT_LAUNCH(1), kern<<<grid, FAST_THREADS, smem, st>>>(p);
  return 1;
}
template <int MODE, class LT>
inline int launch_form3(const Form3Params& p, int rb, size_t smem, cudaStream_t st, int sms) {
  switch (rb) {
    case 4: return launch_form3_rb<MODE, LT, 4>(p, smem, st, sms);
    case 2: return launch_form3_rb<MODE, LT, 2>(p, smem, st, sms);
    default: return launch_form3_rb<MODE, LT, 1>(p, smem, st, sms);
  }
}

// returns 1 if it accumulated into the global tables (which it also initialises), 0 to fall back
inline int try_form3_path(const agc_request_t* r, const Tables& t, cudaStream_t st, int sms) {
  const agc_index_t& ix = r->index;
  if (ix.form != AGC_FORM_AXIS || ix.axis != ix.ndim - 1 || ix.ndim < 2) return 0;
  if (ix.dtype != AGC_DT_I64 && ix.dtype != AGC_DT_I32) return 0;
  const int op = r->op;
  const bool nanskip = r->flags & AGC_F_NANSKIP;
  if (!(op == AGC_OP_SUM || op == AGC_OP_SUMSQ || op == AGC_OP_MEAN || op == AGC_OP_MIN || op == AGC_OP_MAX || op == AGC_OP_VAR))
    return 0;
  if (r->a == nullptr || r->a_dtype != AGC_DT_F32) return 0;
  const long long cols = ix.dims[ix.ndim - 1], G = ix.axis_size;
  if (cols < 4096 || (cols & 3) || G <= 0 || r->n < (1 << 16)) return 0;
  // output must be laid out as [rows][G] (C order): flat = outer * G + label
  if (ix.strides[ix.ndim - 1] != 1) return 0;
  long long expect = G, rows = 1;
  for (int d = ix.ndim - 2; d >= 0; --d) { if (ix.strides[d] != expect) return 0; expect *= ix.dims[d]; rows *= ix.dims[d]; }
  if (rows * cols != r->n || rows * G != r->size) return 0;
  if (((uintptr_t)ix.labels & 15) || ((uintptr_t)r->a & 15)) return 0;
  int mode;
  const bool want_b1 = (op == AGC_OP_SUM || op == AGC_OP_SUMSQ) && (r->flags & AGC_F_NEED_TOUCHED);
  if (op == AGC_OP_MAX) mode = FM_MAX;
  else if (op == AGC_OP_MIN) mode = FM_MIN;
  else if (op == AGC_OP_MEAN || op == AGC_OP_VAR || want_b1) mode = FM_SUMCNT;
  else mode = FM_SUM;
  const int words = mode == FM_SUM ? 2 : (mode == FM_SUMCNT ? 3 : 1);
  // two CTAs per SM: up to 2 x 97 KB the L1 keeps 60 KB for the loads in flight (see AGC_SMEM1 in agc_dispatch.cuh); a single
  // row block may go up to 110 KB (slower streaming, still far ahead of the general kernel)
  const size_t cap = 97 * 1024, cap_max = 110 * 1024;
  int rb = 4;                                              // 8 rows of 128-bit loads would spill at 64 registers / thread
  while (rb > 1 && (size_t)rb * G * words * 4 > cap) rb >>= 1;
  if ((size_t)rb * G * words * 4 > cap_max) return 0;
  while (rb > 1 && rows < rb) rb >>= 1;

  Form3Params p;
  p.labels = ix.labels; p.a = (const float*)r->a; p.rows = rows; p.cols = cols; p.G = G; p.t = t; p.status = r->status;
  p.flags = r->flags; p.square = op == AGC_OP_SUMSQ; p.want_b1 = want_b1; p.want_cnt = op == AGC_OP_MEAN || op == AGC_OP_VAR;
  const long long cvec = cols >> 2;
  const long long n_row_blocks = (rows + rb - 1) / rb;
  // enough items for dynamic balance (>= 8 per CTA) while a chunk stays >= 4 tiles and within the fixed-point headroom
  long long chunks = 1;
  while (n_row_blocks * chunks < (long long)sms * 16 && cvec / (chunks * 2) >= 4 * FAST_THREADS) chunks *= 2;
  while ((cvec + chunks - 1) / chunks * 4 > FAST_MAX_PER_CTA) chunks *= 2;
  p.chunk_vec = (cvec + chunks - 1) / chunks;
  p.n_col_chunks = (int)((cvec + p.chunk_vec - 1) / p.chunk_vec);
  if (n_row_blocks * p.n_col_chunks > 0x7fffffffll) return 0;
  p.n_items = (int)(n_row_blocks * p.n_col_chunks);
  p.counter = r->status + 7;
  cudaMemsetAsync(p.counter, 0, sizeof(int), st);
  AGC_COUNT_LAUNCH(1), k_init_tables<<<(int)((r->size + 255) / 256), 256, 0, st>>>(t, r->size, r->op, r->a_dtype, r->out_dtype, 0);
  const size_t smem = (size_t)rb * G * words * 4;
  const bool i64 = ix.dtype == AGC_DT_I64;
  int rc = 0;
#define AGC_F3_CASE(M) if (mode == M) rc = i64 ? launch_form3<M, long long>(p, rb, smem, st, sms) : launch_form3<M, int>(p, rb, smem, st, sms);
  AGC_F3_CASE(FM_SUM) AGC_F3_CASE(FM_SUMCNT) AGC_F3_CASE(FM_MAX) AGC_F3_CASE(FM_MIN)
#undef AGC_F3_CASE
  (void)nanskip;
  return rc;
}

}  // namespace agc
