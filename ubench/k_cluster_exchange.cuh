// agc_cluster.cuh -- group tables too large for one SM's shared memory (28 K < size <~ 4*10^5):
// the table is PARTITIONED over the CTAs of a thread-block cluster (group g lives in CTA g % C at
// slot g / C) and the elements are exchanged between the CTAs in bulk.
//
// Why not remote atomics: red.shared::cluster is packet-bound (~45-90 G op/s on the whole chip,
// profiles/r01_probe2_fixedpoint_cluster.log) and global RED tops out at ~170 Gelem/s.  Coalesced
// DSMEM *reads* are bytes-bound (17-21 B/clk/SM), so each phase is
//   1. every CTA takes a tile of 4096 (label, value) pairs (128-bit streaming loads, prefetched into
//      registers one phase ahead) and classifies them: destination CTA, slot, payload,
//   2. counting-sorts the tile by destination into an outbox in its own shared memory -- no atomics:
//      a packed per-lane histogram (8 destinations x 8 bit), one warp scan, a REDUX-based block scan;
//      the order inside a destination is the element order, so the result is deterministic for a
//      given launch shape,
//   3. __syncthreads + barrier.cluster.arrive (relaxed) ... classify the NEXT tile ... barrier.cluster.wait,
//   4. every CTA PULLS the entries addressed to it from the outboxes of all C CTAs
//      (ld.shared::cluster.v2, coalesced, 8 B/entry) and applies them to its private table with
//      native 32-bit ATOMS exactly like k_fast.
// Two outbox buffers and one cluster barrier per phase (the buffer written in phase k+2 was last
// read in phase k, which every CTA finished before it arrived at barrier k+1).
// Accumulator formats, the exact fixed-point window and the bypass are those of agc_fast.cuh; the
// reference exponent is agreed once per cluster.
// NOT PART OF THE LIBRARY.  Kept as the record of a measured experiment: correct in
// hundreds of parity runs but instruction-issue bound (94-124 Gelem/s at 10^5 groups), and its relaxed cluster
// barrier made it intermittently unsafe (one sporadic illegal instruction with `.aligned` barriers, one sporadic
// wrong result without) -- the __syncthreads() + barrier.cluster.arrive.relaxed publication of shared memory is not
// a substitute for arrive.release.  Needs agc_fast.cuh for FastParams / fx_add if it is ever revived.
#pragma once
#include "../numpy_groupies_b200/csrc/agc_fast.cuh"

namespace agc {


#ifndef AGC_CL_THREADS
#define AGC_CL_THREADS 512
#endif
constexpr int CL_THREADS = AGC_CL_THREADS;
constexpr int CL_WARPS = CL_THREADS / 32;
constexpr int CL_TILE = 4096;                       // entries per outbox buffer
constexpr int CL_EPT = CL_TILE / CL_THREADS;        // elements per thread per phase (8; <= 8: 8-bit lane histograms)
constexpr int CL_VEC = CL_EPT / 4;                  // 128-bit value vectors per thread per phase
constexpr int CL_NBUF = 2;
constexpr int CL_OFFS = 16;                         // ints per offset record (C + 1 <= 9 used)
static_assert(CL_EPT == 4 || CL_EPT == 8, "tile shape");

struct ClusterParams {
  FastParams f;
  int logc;            // cluster size C = 1 << logc  (2, 4 or 8)
  int slots;           // table slots per CTA (multiple of 4, >= ceil(size / C))
  int phases;          // tiles per CTA (identical for every CTA: all of them take every barrier)
  int flush_every;     // phases between in-kernel merges (fixed-point / u32-count headroom)
  unsigned covered;    // groups [0, covered) live in the cluster's tables, the rest go straight to the global tables
  const int* choose;   // device word written by k_skew_probe; the kernel runs iff *choose == want
  int want;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ unsigned ld_cluster_u32(unsigned addr) {
  unsigned v; asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_cluster_u16(unsigned addr) {
  unsigned short v; asm volatile("ld.shared::cluster.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory"); return v;
}
__device__ __forceinline__ uint2 ld_cluster_v2(unsigned addr) {
  uint2 v; asm volatile("ld.shared::cluster.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory"); return v;
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { __syncwarp(); asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
// (none of the cluster barriers uses `.aligned`: the classify step ends in data-dependent branches and an aligned
// barrier issued by a warp the compiler has not reconverged is undefined -- it showed up as a sporadic
// 'illegal instruction'.)
// relaxed arrive: no MEMBAR.ALL.GPU (which would also wait for the prefetched global loads of the next tile).
// The outbox stores are ordered by the __syncthreads() that precedes it: BAR.SYNC drains the CTA's pending STS
// into the SM's shared memory, which is the only copy a peer's ld.shared::cluster can observe.
__device__ __forceinline__ void cluster_arrive_relaxed() { __syncwarp(); asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { __syncwarp(); asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }

// 4 x 8-bit fields -> 4 x 16-bit fields
__device__ __forceinline__ unsigned long long spread8to16(unsigned v) {
  unsigned long long x = v;
  x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
  x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
  return x;
}

// ---------------------------------------------------------------------------------------------
template <int MODE, class LT, bool LOAD_A, int LOGC>
__global__ void __launch_bounds__(CL_THREADS, 1) k_cluster(const ClusterParams cp) {
  if (cp.choose && *cp.choose != cp.want) return;        // uniform over the grid: nobody reaches a barrier
  const FastParams& p = cp.f;
  extern __shared__ __align__(16) unsigned sw[];
  __shared__ unsigned s_maxbits;
  constexpr int WORDS = MODE == FM_SUM ? 2 : (MODE == FM_SUMCNT ? 3 : 1);
  constexpr bool IS_SUM = MODE == FM_SUM || MODE == FM_SUMCNT;
  constexpr int C = 1 << LOGC;
  const int S = cp.slots;
  unsigned* w0 = sw;             // sum.lo | key | count (FM_LEN)
  unsigned* w1 = sw + S;         // sum.hi
  unsigned* w2 = sw + 2 * S;     // count (FM_SUMCNT)
  uint2* ent = (uint2*)(sw + WORDS * S);                                  // [NBUF][TILE] (slot, payload)
  int* off = (int*)(ent + CL_NBUF * CL_TILE);                             // [NBUF][CL_OFFS] segment starts per destination
  uint4* wtot = (uint4*)(off + CL_NBUF * CL_OFFS);                        // [WARPS] per-warp totals, 8 x 16 bit
  unsigned short* wbase = (unsigned short*)(wtot + CL_WARPS);             // [WARPS][16] outbox position of (warp, destination)
  const unsigned init0 = MODE == FM_MAX ? 0x80000000u : (MODE == FM_MIN ? 0x7fffffffu : 0u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned rank = cluster_ctarank();
  const unsigned size32 = (unsigned)p.size, covered = cp.covered;

  for (int i = tid; i < WORDS * S; i += CL_THREADS) sw[i] = (i < S) ? init0 : 0u;
  if (tid == 0) s_maxbits = 0u;
  __syncthreads();

  const long long nvec = p.n >> 2;
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  const bool square = p.square;
  bool bad = false;

  // ---- raw registers of the tile in flight ----------------------------------------------------------
  int4 rl[CL_VEC][2], ra[CL_VEC]; bool have[CL_VEC];
  auto fetch = [&](int phase) {
#pragma unroll
    for (int j = 0; j < CL_VEC; ++j) {
      const long long v = (((long long)blockIdx.x + (long long)phase * gridDim.x) * CL_VEC + j) * CL_THREADS + tid;
      have[j] = phase < cp.phases && v < nvec;
      rl[j][0] = make_int4(0, 0, 0, 0); rl[j][1] = rl[j][0]; ra[j] = rl[j][0];
      if (have[j]) {
        if (sizeof(LT) == 8) { rl[j][0] = ldg_stream16((const int4*)p.labels + 2 * v); rl[j][1] = ldg_stream16((const int4*)p.labels + 2 * v + 1); }
        else rl[j][0] = ldg_stream16((const int4*)p.labels + v);
        if (LOAD_A) ra[j] = ldg_stream16((const int4*)p.a + v);
      }
    }
  };
  fetch(0);

  // ---- reference exponent, agreed over the cluster ------------------------------------------------
  unsigned lo_bits = 0, span_bits = 0; double scale = 1.0, inv_scale = 1.0;
  if (IS_SUM) {
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < CL_VEC; ++j) {
      if (!have[j]) continue;
      unsigned b[4] = {(unsigned)ra[j].x & 0x7fffffffu, (unsigned)ra[j].y & 0x7fffffffu, (unsigned)ra[j].z & 0x7fffffffu, (unsigned)ra[j].w & 0x7fffffffu};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned bb = b[k];
        if (square) { float f = __uint_as_float(bb); bb = __float_as_uint(f * f); }
        if (bb < 0x7f800000u && bb > m) m = bb;
      }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0 && m) atomicMax(&s_maxbits, m);
  }
  cluster_arrive(); cluster_wait();                       // tables initialised, s_maxbits final everywhere
  if (IS_SUM) {
    unsigned mb = 0;
    const unsigned a_mb = smem_u32(&s_maxbits);
    for (int r = 0; r < C; ++r) mb = max(mb, ld_cluster_u32(mapa_u32(a_mb, r)));
    int eb = (int)(mb >> 23);
    if (eb == 0) eb = 127;
    eb += 1;
    int lo_e = eb - FX_WINDOW; if (lo_e < 1) lo_e = 1;
    lo_bits = (unsigned)lo_e << 23;
    span_bits = (unsigned)(eb - lo_e) << 23;
    scale = __longlong_as_double((long long)(1023 + FX_F - (eb - 127)) << 52);
    inv_scale = __longlong_as_double((long long)(1023 - FX_F + (eb - 127)) << 52);
  }

  // merge the CTA's table into the global accumulators (one RED per touched group), optionally reset
  auto flush = [&](bool reset) {
    for (int s = tid; s < S; s += CL_THREADS) {
      const long long gi = ((long long)s << LOGC) | rank;
      if (gi >= (long long)covered) continue;
      if (MODE == FM_LEN) {
        unsigned c = w0[s];
        if (c) { atomicAdd((unsigned long long*)(p.t.t1 + gi), (unsigned long long)c); if (reset) w0[s] = 0u; }
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
        if (q) { atomicAdd(p.t.t0f + gi, (double)q * inv_scale); if (reset) { w0[s] = 0u; w1[s] = 0u; } }
        if (MODE == FM_SUMCNT) {
          unsigned c = w2[s];
          if (c) {
            if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + gi), (unsigned long long)c);
            if (p.want_b1) p.t.b1[gi] = 1;
            if (reset) w2[s] = 0u;
          }
        }
      }
    }
  };

  // ---- classified tile: lab = label (0xffffffff: nothing to forward), py = payload -------------------
  unsigned lab[CL_EPT], py[CL_EPT];
  unsigned lex0 = 0, lex1 = 0;                            // this lane's exclusive prefix per destination, 8-bit fields
  auto classify = [&]() {
    unsigned h0 = 0, h1 = 0;                              // lane histogram: destinations 0-3 | 4-7, 8-bit fields
#pragma unroll
    for (int j = 0; j < CL_VEC; ++j) {
      unsigned glo[4], ghi[4];
      if (sizeof(LT) == 8) {
        glo[0] = rl[j][0].x; ghi[0] = rl[j][0].y; glo[1] = rl[j][0].z; ghi[1] = rl[j][0].w;
        glo[2] = rl[j][1].x; ghi[2] = rl[j][1].y; glo[3] = rl[j][1].z; ghi[3] = rl[j][1].w;
      } else {
        glo[0] = rl[j][0].x; glo[1] = rl[j][0].y; glo[2] = rl[j][0].z; glo[3] = rl[j][0].w;
        ghi[0] = ghi[1] = ghi[2] = ghi[3] = 0u;           // a negative int32 label fails glo < size32 (size < 2^31)
      }
      const float xs[4] = {__int_as_float(ra[j].x), __int_as_float(ra[j].y), __int_as_float(ra[j].z), __int_as_float(ra[j].w)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = j * 4 + e;
        const bool valid = ghi[e] == 0u && glo[e] < size32;
        bad |= have[j] && !valid;
        bool fwd = have[j] && valid;
        float xv = xs[e];
        unsigned pv = 0u;
        if (fwd && glo[e] >= covered) {                      // partial coverage: this group's accumulator is the global one
          fwd = false;
          const bool nan = xv != xv;
          if (!(LOAD_A && nanskip && nan)) {
            if (MODE == FM_LEN) atomicAdd((unsigned long long*)(p.t.t1 + glo[e]), 1ull);
            else if (MODE == FM_MAX || MODE == FM_MIN) {
              const long long k64 = nan ? (MODE == FM_MAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX) : key_of((double)xv);
              if (MODE == FM_MAX) atomicMax(p.t.t0i + glo[e], k64); else atomicMin(p.t.t0i + glo[e], k64);
            } else {
              const float x2 = square ? xv * xv : xv;
              if ((__float_as_uint(x2) & 0x7fffffffu) != 0u) atomicAdd(p.t.t0f + glo[e], (double)x2);
              if (MODE == FM_SUMCNT) { if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + glo[e]), 1ull); if (p.want_b1) p.t.b1[glo[e]] = 1; }
            }
          }
        }
        if (MODE == FM_LEN) {
          if (LOAD_A) fwd = fwd && !(nanskip && xv != xv);
        } else if (MODE == FM_MAX || MODE == FM_MIN) {
          const bool nan = xv != xv;
          fwd = fwd && !(nanskip && nan);
          pv = nan ? (MODE == FM_MAX ? 0x7fffffffu : 0x80000000u) : (unsigned)key32_of(xv);   // NaN wins
        } else {
          const bool nan = xv != xv;
          fwd = fwd && !(nanskip && nan);
          if (square) xv = xv * xv;
          const unsigned ab = __float_as_uint(xv) & 0x7fffffffu;
          const bool inwin = ab - lo_bits < span_bits;
          pv = inwin ? __float_as_uint(xv) : 0u;
          if (fwd && !inwin) {                             // zero, or out of window (huge / tiny / inf / NaN): exact f64 RED
            if (ab != 0u) atomicAdd(p.t.t0f + glo[e], (double)xv);
            if (MODE == FM_SUM) fwd = false;               // FM_SUMCNT still forwards +0.0 for the count
          }
        }
        lab[i] = fwd ? glo[e] : 0xffffffffu;
        py[i] = pv;
        if (fwd) {
          const unsigned inc = 1u << ((glo[e] & 3u) << 3);
          if (LOGC == 3 && (glo[e] & 4u)) h1 += inc; else h0 += inc;
        }
      }
    }
    // inclusive warp scan of the packed histograms; a field that reaches 256 wraps, the exclusive value stays right
    unsigned i0 = h0, i1 = h1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t0 = __shfl_up_sync(0xffffffffu, i0, o);
      if (lane >= o) i0 += t0;
      if (LOGC == 3) { const unsigned t1 = __shfl_up_sync(0xffffffffu, i1, o); if (lane >= o) i1 += t1; }
    }
    lex0 = i0 - h0; lex1 = i1 - h1;
    if (lane == 31) {                                     // warp totals, widened to 16-bit fields
      uint4 t;
      t.x = __byte_perm(lex0, 0, 0x4140) + __byte_perm(h0, 0, 0x4140);
      t.y = __byte_perm(lex0, 0, 0x4342) + __byte_perm(h0, 0, 0x4342);
      t.z = __byte_perm(lex1, 0, 0x4140) + __byte_perm(h1, 0, 0x4140);
      t.w = __byte_perm(lex1, 0, 0x4342) + __byte_perm(h1, 0, 0x4342);
      wtot[warp] = t;
    }
  };
  // positions of (warp, destination) in the outbox: REDUX over the per-warp totals, destination-major order
  int tile_total = 0;
  auto block_scan = [&]() {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (lane < CL_WARPS) v = wtot[lane];
    const bool before = lane < warp;
    const unsigned p0 = __reduce_add_sync(0xffffffffu, before ? v.x : 0u), p1 = __reduce_add_sync(0xffffffffu, before ? v.y : 0u);
    const unsigned t0 = __reduce_add_sync(0xffffffffu, v.x), t1 = __reduce_add_sync(0xffffffffu, v.y);
    unsigned p2 = 0, p3 = 0, t2 = 0, t3 = 0;
    if (LOGC == 3) {
      p2 = __reduce_add_sync(0xffffffffu, before ? v.z : 0u); p3 = __reduce_add_sync(0xffffffffu, before ? v.w : 0u);
      t2 = __reduce_add_sync(0xffffffffu, v.z); t3 = __reduce_add_sync(0xffffffffu, v.w);
    }
    // exclusive prefix over the destinations (16-bit fields; every partial sum <= 4096)
    const unsigned long long ta = ((unsigned long long)t1 << 32) | t0, tb = ((unsigned long long)t3 << 32) | t2;
    const unsigned long long ia = ta * 0x0001000100010001ull;
    const unsigned long long ib = (tb + (ia >> 48)) * 0x0001000100010001ull;
    const unsigned long long ea = ia - ta + (((unsigned long long)p1 << 32) | p0);
    const unsigned long long eb = ib - tb + (((unsigned long long)p3 << 32) | p2);
    tile_total = (int)(LOGC == 3 ? (ib >> 48) : (ia >> 48));
    if (lane == 0) {
      uint4 o; o.x = (unsigned)ea; o.y = (unsigned)(ea >> 32); o.z = (unsigned)eb; o.w = (unsigned)(eb >> 32);
      *(uint4*)(wbase + warp * 16) = o;
    }
    __syncwarp();
  };
  auto scatter = [&](int buf) {
    uint2* o = ent + buf * CL_TILE;
    const unsigned short* wb = wbase + warp * 16;
    unsigned r0 = lex0, r1 = lex1;
#pragma unroll
    for (int i = 0; i < CL_EPT; ++i) {
      const unsigned g = lab[i];
      if (g == 0xffffffffu) continue;
      const unsigned d = g & (unsigned)(C - 1);
      const unsigned sh = (g & 3u) << 3;
      const bool hi = LOGC == 3 && (g & 4u);
      const unsigned pos = (unsigned)wb[d] + (((hi ? r1 : r0) >> sh) & 0xffu);
      if (hi) r1 += 1u << sh; else r0 += 1u << sh;
      o[pos] = make_uint2(g >> LOGC, py[i]);
    }
    if (warp == 0 && lane <= C) off[buf * CL_OFFS + lane] = lane < C ? (int)wb[lane] : tile_total;   // warp 0: positions = segment starts
  };

  // pull the entries addressed to this CTA out of every CTA's outbox `buf` and apply them
  constexpr int PER = CL_THREADS >> LOGC;                 // threads serving one source CTA
  const unsigned src = (unsigned)tid / PER, sub = (unsigned)tid % PER;
  auto consume = [&](int buf) {
    const unsigned a_off = mapa_u32(smem_u32(off + buf * CL_OFFS), src);
    const unsigned a_ent = mapa_u32(smem_u32(ent + buf * CL_TILE), src);
    const int lo = (int)ld_cluster_u32(a_off + rank * 4), hi = (int)ld_cluster_u32(a_off + rank * 4 + 4);
    for (int base = lo + (int)sub; base < hi; base += 8 * PER) {
      uint2 en[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * PER;
        en[u] = make_uint2(0xffffffffu, 0u);
        if (i < hi) en[u] = ld_cluster_v2(a_ent + 8u * i);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const unsigned s = en[u].x;
        if (s == 0xffffffffu) continue;
        if (MODE == FM_LEN) atomicAdd(w0 + s, 1u);
        else if (MODE == FM_MAX) atomicMax((int*)w0 + s, (int)en[u].y);
        else if (MODE == FM_MIN) atomicMin((int*)w0 + s, (int)en[u].y);
        else {
          const long long q = __double2ll_rn((double)__uint_as_float(en[u].y) * scale);
          if (q) fx_add(w0 + s, w1 + s, q);
          if (MODE == FM_SUMCNT) atomicAdd(w2 + s, 1u);
        }
      }
    }
  };

  // ---- software pipeline ------------------------------------------------------------------------------
  if (cp.phases > 0) {
    classify(); fetch(1);
    __syncthreads();
    block_scan();
  }
  for (int k = 0; k < cp.phases; ++k) {
    const int buf = k & 1;
    scatter(buf);
    __syncthreads();                                        // outbox k complete in this SM's shared memory
    cluster_arrive_relaxed();                               // outbox k published
    const bool more = k + 1 < cp.phases;
    if (more) { classify(); fetch(k + 2); }                 // the next tile is classified under the barrier latency
    cluster_wait();
    consume(buf);
    if (MODE != FM_MAX && MODE != FM_MIN && (k + 1) % cp.flush_every == 0 && more) {
      __syncthreads();                                      // every thread of this CTA is done with consume(k)
      flush(true);
    }
    __syncthreads();                                        // wtot of tile k+1 complete
    if (more) block_scan();
  }
  // tail elements (n % 4) go straight to the global tables
  if (blockIdx.x == 0 && tid < (int)(p.n & 3)) {
    const long long i = (nvec << 2) + tid;
    const long long gg = LabelVec<LT>::one(p.labels, i);
    float xv = LOAD_A ? p.a[i] : 0.f;
    if ((unsigned long long)gg >= (unsigned long long)p.size) bad = true;
    else if (!(LOAD_A && nanskip && xv != xv)) {
      if (MODE == FM_LEN) atomicAdd((unsigned long long*)(p.t.t1 + gg), 1ull);
      else if (MODE == FM_MAX || MODE == FM_MIN) {
        const bool nan = xv != xv;
        long long k64 = nan ? (MODE == FM_MAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX) : key_of((double)xv);
        if (MODE == FM_MAX) atomicMax(p.t.t0i + gg, k64); else atomicMin(p.t.t0i + gg, k64);
      } else {
        if (square) xv = xv * xv;
        if ((__float_as_uint(xv) & 0x7fffffffu) != 0u) atomicAdd(p.t.t0f + gg, (double)xv);
        if (MODE == FM_SUMCNT) { if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + gg), 1ull); if (p.want_b1) p.t.b1[gg] = 1; }
      }
    }
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && tid == 0) p.status[AGC_ST_REGIME] = 30 + MODE;
  cluster_arrive(); cluster_wait();                         // nobody leaves while a peer still reads its outbox
  flush(false);
}

inline size_t cluster_smem_bytes(int mode, int slots) {
  const int words = mode == FM_SUM ? 2 : (mode == FM_SUMCNT ? 3 : 1);
  return (size_t)words * slots * 4 + (size_t)CL_NBUF * CL_TILE * 8 + CL_NBUF * CL_OFFS * 4 + CL_WARPS * 16 + CL_WARPS * 32;
}

template <int MODE, class LT, bool LOAD_A, int LOGC>
inline int launch_cluster_c(ClusterParams cp, size_t smem, cudaStream_t st) {
  auto kern = k_cluster<MODE, LT, LOAD_A, LOGC>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    attr_smem = smem;
  }
  constexpr int C = 1 << LOGC;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  // how many clusters fit on the device at once (GPC packing: C = 4 / 8 strand a few SMs)
  static int cached_clusters = 0;
  static size_t cached_smem = 0;
  if (cached_clusters == 0 || cached_smem != smem) {
    cfg.gridDim = dim3(C);
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) != cudaSuccess || ncl <= 0) { cudaGetLastError(); return 0; }
    cached_clusters = ncl; cached_smem = smem;
  }
  const int grid = cached_clusters * C;
  const long long nvec = cp.f.n >> 2;
  const long long tiles = (nvec + CL_TILE / 4 - 1) / (CL_TILE / 4);
  cp.phases = (int)((tiles + grid - 1) / grid);
  // a CTA's table receives the elements of its whole cluster: merge every 2^22 elements (see FAST_MAX_PER_CTA)
  cp.flush_every = (int)(FAST_MAX_PER_CTA / ((long long)C * CL_TILE));
  cfg.gridDim = dim3(grid);
  if (cudaLaunchKernelEx(&cfg, kern, cp) != cudaSuccess) return -31;
  AGC_COUNT_LAUNCH(1);
  return 1;
}
template <int MODE, class LT, bool LOAD_A>
inline int launch_cluster(const ClusterParams& cp, size_t smem, cudaStream_t st) {
  return cp.logc == 3 ? launch_cluster_c<MODE, LT, LOAD_A, 3>(cp, smem, st) : launch_cluster_c<MODE, LT, LOAD_A, 2>(cp, smem, st);
}

// picks the cluster size for (mode, size); returns 0 if the table does not fit an 8-CTA cluster
inline int cluster_plan(int mode, long long size, int* logc, int* slots, size_t* smem) {
  const size_t cap = 227 * 1024 - 64;
  for (int lc = 2; lc <= 3; ++lc) {
    const long long C = 1ll << lc;
    long long s = (size + C - 1) / C;
    s = (s + 3) & ~3ll;
    const size_t need = cluster_smem_bytes(mode, (int)s);
    if (need <= cap) { *logc = lc; *slots = (int)s; *smem = need; return 1; }
  }
  return 0;
}

}  // namespace agc
