// agc_team.cuh -- k_team: sums (and counts) for tables that do not fit ONE SM, with no scattered global RED at all.
//
// The per-SM budget is ~195 KB of shared memory (agc_dispatch.cuh: more costs a third of the load bandwidth), i.e.
// 49.9 K one-word sums.  At 10^5 groups the partial-table kernel therefore sends ~half of the elements to the global
// table as RED packets and is bound by them (0.38 RED/clk/SM, round-1 profile: 4.06 ms where the stream needs 1.97).
// Here a thread-block CLUSTER of C SMs owns the table together: group g lives in CTA (g mod C) at slot g / C.  A
// warp streams its share of (labels, a) like k_fast; an element of a group it owns goes straight into its table
// (native ATOMS.ADD, one-word fixed point with carry forwarding, see FM_SUM1); every other element becomes an 8-byte
// record {q, slot} pushed into the owner's shared memory over DSMEM with `st.async` (completion counted by an mbarrier
// in the owner's shared memory -- no fence, no cluster barrier in the loop).  Warp w of every CTA talks only to warp w
// of its peers: per (peer, warp) two record buffers, a "full" mbarrier at the receiver (one arrive.expect_tx per
// phase, posted remotely by the sender) and an "empty" mbarrier at the sender (posted remotely by the receiver once
// it has applied the records).  A warp applies what it received for phase p-1 right after it has sent phase p, so the
// round trip hides behind its own streaming.  Records that do not fit their buffer (more than mean + 2 sigma of a
// phase went to one peer) and groups without a slot (cov < ceil(size / C)) fall back to the global RED.
//
// Exactness is FM_SUM1's: in-window values convert exactly (6 binades), carries out of the 32-bit word are forwarded at
// once as exact f64 REDs of +-2^32 units, everything outside the window (huge / tiny / inf / NaN) is one f64 RED.
#pragma once
#include "agc_fast.cuh"

namespace agc {

constexpr int TEAM_THREADS = 1024;
constexpr int TEAM_WARPS = 32;
enum { TM_SUM = 0, TM_SUMCNT = 1 };

struct TeamParams {
  const void* labels; const float* a;
  long long n, size;
  Tables t;
  int* status;
  int flags;            // AGC_F_NANSKIP
  int want_cnt;         // TM_SUMCNT: merge counts into T1
  int want_b1;          // TM_SUMCNT: write touched flags
  int cov;              // table slots per CTA (group g -> CTA g % C, slot g / C; slots >= cov go to the global tables)
  int cap;              // 8-byte record slots per (peer, buffer, warp), slot 0 is the header {count}
  const int* choose;    // device word written by k_skew_probe; run iff *choose == want (NULL: always)
  int want;
};

namespace dsm {
// a value the compiler must keep in a register instead of re-deriving it at every use (the hot loop is instruction bound)
__device__ __forceinline__ unsigned keep(unsigned x) { unsigned y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned mapa(unsigned saddr, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// all threads of the cluster (not .aligned: round 1 saw sporadic faults with aligned cluster barriers behind divergent code)
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) { while (!mbar_try_wait(bar, parity)) {} }
// 8-byte record into a peer's shared memory; its arrival is counted (8 bytes) by the peer's mbarrier
__device__ __forceinline__ void st_async_rec(unsigned raddr, unsigned lo, unsigned hi, unsigned rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
               :: "r"(raddr), "r"(lo), "r"(hi), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_remote_rec(unsigned raddr, unsigned lo, unsigned hi) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" :: "r"(raddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void st_remote_u32(unsigned raddr, unsigned v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(raddr), "r"(v) : "memory");
}
// one arrival on a peer's mbarrier that also announces `bytes` of st.async traffic (may already have landed)
__device__ __forceinline__ void remote_arrive_expect_tx(unsigned rbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" :: "r"(rbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void remote_arrive_relaxed(unsigned rbar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" :: "r"(rbar) : "memory");
}
__device__ __forceinline__ void remote_arrive_release(unsigned rbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(rbar) : "memory");
}
}  // namespace dsm

// Instruction budget: the kernel issues ~1.7 warp instructions per element (C = 2), so everything that is not on the
// common path -- labels outside [0, cov * C), values outside the fixed-point window, NaN handling, a full buffer --
// sits behind ONE rarely taken branch per element (`slow`).
template <int C, class LT, int MODE, bool NANSKIP>
__global__ void __launch_bounds__(TEAM_THREADS, 1) k_team(const TeamParams p) {
  if (p.choose && *p.choose != p.want) return;              // same decision in every CTA of the cluster
  extern __shared__ __align__(16) unsigned sw[];
  __shared__ unsigned s_maxbits;
  __shared__ unsigned s_peer_max[8];
  constexpr int LOGC = C == 2 ? 1 : (C == 4 ? 2 : 3);
  constexpr int P = C - 1;                                   // peers
  constexpr bool HAS_CNT = MODE == TM_SUMCNT;
  static_assert(C == 2 || C == 4 || C == 8, "cluster size");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned me = dsm::ctarank();
  const int cov = p.cov, cap = p.cap;

  // ---- shared-memory layout: [sum words | 16-bit counters] [inbox records] [mbarriers] ------------------------
  unsigned* table = sw;
  unsigned* cnt16 = sw + cov;
  const int table_words = cov + (HAS_CNT ? (cov + 1) / 2 : 0);
  const int inbox_off = (table_words + 3) & ~3;
  unsigned long long* inbox = (unsigned long long*)(sw + inbox_off);               // [(j * 2 + b) * 32 + w][cap]
  unsigned long long* bars = inbox + (size_t)P * 2 * TEAM_WARPS * cap;             // full[P][2][32], empty[P][2][32]
  for (int i = threadIdx.x; i < table_words; i += TEAM_THREADS) sw[i] = i < cov ? 0x80000000u : 0u;
  for (int i = threadIdx.x; i < 2 * P * 2 * TEAM_WARPS; i += TEAM_THREADS) dsm::mbar_init(dsm::smem_u32(bars + i), 1);
  if (threadIdx.x == 0) s_maxbits = 0u;
  dsm::fence_mbar_init();
  dsm::cluster_sync();                // every CTA of the cluster is running and has initialised its barriers: DSMEM may be touched

  const long long nvec = p.n >> 2;
  const long long grid_vecs = (long long)gridDim.x * TEAM_THREADS;
  const long long nphase = (nvec + grid_vecs - 1) / grid_vecs;                     // the same for every warp of the grid
  long long v = (long long)blockIdx.x * TEAM_THREADS + threadIdx.x;
  int bad = 0;

  // ---- one reference exponent per cluster (records carry fixed-point integers) ----------------------------------
  {
    unsigned m = 0;
    if (v < nvec) {
      const int4 av = ldg_stream16((const int4*)p.a + v);
      const unsigned b[4] = {(unsigned)av.x & 0x7fffffffu, (unsigned)av.y & 0x7fffffffu, (unsigned)av.z & 0x7fffffffu, (unsigned)av.w & 0x7fffffffu};
#pragma unroll
      for (int k = 0; k < 4; ++k) if (b[k] < 0x7f800000u && b[k] > m) m = b[k];
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0 && m) atomicMax(&s_maxbits, m);
    __syncthreads();
    if (threadIdx.x < C) dsm::st_remote_u32(dsm::mapa(dsm::smem_u32(&s_peer_max[me]), threadIdx.x), s_maxbits);
  }
  dsm::cluster_sync();                // the exponent words are visible cluster-wide behind this
  unsigned lo_bits = 0, span_bits = 0; float scale_f = 1.f; double inv_scale = 1.0;
  {
    unsigned mb = 0;
#pragma unroll
    for (int r = 0; r < C; ++r) mb = max(mb, s_peer_max[r]);
    int eb = (int)(mb >> 23);
    if (eb == 0) eb = 127;
    eb += 1;
    int lo_e = eb - FX1_WINDOW; if (lo_e < 1) lo_e = 1;
    lo_bits = (unsigned)lo_e << 23;
    span_bits = (unsigned)(eb - lo_e) << 23;
    const int se = FX1_F - (eb - 127);                         // q = x * 2^se, exact for in-window x
    if (se < -126 || se > 127) span_bits = 0;                  // scale not a normal float: nothing converts, all bypass
    else scale_f = __int_as_float((se + 127) << 23);
    inv_scale = __longlong_as_double((long long)(1023 - se) << 52);
  }
  const unsigned cnt_sbase = dsm::keep(dsm::smem_u32(cnt16));
  const unsigned table_s = dsm::keep(dsm::smem_u32(table));
  const unsigned cov_c = (unsigned)cov << LOGC;               // labels below this have a slot somewhere in the cluster (host: cov * C <= size)
  const unsigned lt_mask = dsm::keep((1u << lane) - 1u);
  const unsigned capm = (unsigned)cap;

  // ---- per-peer addresses: peer j = CTA (me + j + 1) % C receives into ITS region j; I receive region j from CTA (me - j - 1) % C,
  //      which is peer P - 1 - j.  The shared::cluster window of a CTA is linear: remote address = peer base + local offset.
  constexpr unsigned BAR_B = TEAM_WARPS * 8;                   // bytes from a buffer-0 barrier to its buffer-1 twin
  const unsigned buf_stride = (unsigned)(TEAM_WARPS * cap * 8);  // bytes from buffer 0 to buffer 1 of a region
  const unsigned sw_s = dsm::smem_u32(sw);
  const unsigned in0 = dsm::smem_u32(inbox) + (unsigned)warp * cap * 8u;      // my region 0, buffer 0 (region j: + j * 2 * buf_stride)
  const unsigned full0 = dsm::keep(dsm::smem_u32(bars) + (unsigned)warp * 8u);  // full[0][0][warp] (region j: + j * 2 * BAR_B)
  const unsigned empty0 = full0 + P * 2 * BAR_B;                             // empty[0][0][warp]
  unsigned dst0[P], dbar0[P], rempty0[P];                      // buffer-0 addresses at the peers: records, "full" barrier; "empty" barrier of the sender of region j
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const unsigned rb = dsm::mapa(sw_s, (me + j + 1) & (C - 1)) - sw_s;
    const unsigned rs = dsm::mapa(sw_s, (me - j - 1) & (C - 1)) - sw_s;
    dst0[j] = dsm::keep(rb + in0 + (j * 2) * buf_stride);
    dbar0[j] = dsm::keep(rb + full0 + (j * 2) * BAR_B);
    rempty0[j] = dsm::keep(rs + empty0 + (j * 2) * BAR_B);
  }

  // ---- element sinks ---------------------------------------------------------------------------------------
  // (rarely used constants -- the L2 policy, the value of a carry -- are re-derived where they are needed: registers are short)
  const double hi_unit = inv_scale * 4294967296.0;            // what one carry out of a 32-bit word is worth
  auto repair = [&](unsigned wi, int sh, long long g) {        // k_len16's read-check-repair: take 0x8000 out, forward it
    unsigned o = *(volatile unsigned*)(cnt16 + wi);
    while (((o >> sh) & 0xffffu) >= 0x8000u) {
      const unsigned seen = atomicCAS(cnt16 + wi, o, o - (0x8000u << sh));
      if (seen == o) { red_add_u64_hint((unsigned long long*)(p.t.t1 + g), 0x8000ull, l2_evict_last_policy()); break; }
      o = seen;
    }
  };
  // a record (or an own element) into MY table: one native ATOMS.ADD; a carry / borrow out of the 32-bit word (the add
  // wrapped although q >= 0, or did not although q < 0) is forwarded at once as an exact f64 RED of +-2^32 units.
  // Issue and check are separate steps: a warp issues the ATOMS of a whole batch before it looks at the first old value
  // (the kernel is bound by the latency of its dependent steps, not by issue slots).
  auto apply_issue = [&](unsigned li, int q) -> unsigned {
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(table_s + li * 4u), "r"((unsigned)q) : "memory");
    return old;
  };
  auto apply_check = [&](unsigned li, int q, unsigned old) {
    if ((old + (unsigned)q < old) != (q < 0)) atomicAdd(p.t.t0f + ((li << LOGC) | me), q >= 0 ? hi_unit : -hi_unit);
  };
  auto count_one = [&](unsigned li) {
    const int sh = (int)(li & 1u) << 4; const unsigned wi = li >> 1;
    if (((*(volatile unsigned*)(cnt16 + wi) >> sh) & 0xffffu) >= 0x8000u) repair(wi, sh, ((long long)li << LOGC) | me);
    reds_add_u32(cnt_sbase + wi * 4, 1u << sh);
  };
  auto global_count = [&](long long g) {
    if (p.want_cnt) red_add_u64_hint((unsigned long long*)(p.t.t1 + g), 1ull, l2_evict_last_policy());
    if (p.want_b1) p.t.b1[g] = 1;
  };
  auto consume = [&](unsigned ph) {                            // apply what the peers sent me in phase ph
    const unsigned b = ph & 1u, par = (ph >> 1) & 1u;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      dsm::mbar_wait(full0 + (j * 2 + b) * BAR_B, par);
      const unsigned long long* reg = inbox + ((size_t)(j * 2 + b) * TEAM_WARPS + warp) * cap;
      const unsigned cnt = (unsigned)reg[0];
      {
        constexpr int ROUNDS = 3;                              // cap <= 97 records per buffer (host side)
        unsigned long long rec[ROUNDS]; unsigned old[ROUNDS];
#pragma unroll
        for (int t = 0; t < ROUNDS; ++t) { const unsigned i = lane + 32u * t; rec[t] = i < cnt ? reg[1 + i] : 0ull; }
#pragma unroll
        for (int t = 0; t < ROUNDS; ++t) { old[t] = 0u; if ((unsigned)rec[t] != 0u) old[t] = apply_issue((unsigned)(rec[t] >> 32), (int)(unsigned)rec[t]); }
#pragma unroll
        for (int t = 0; t < ROUNDS; ++t) {
          if ((unsigned)rec[t] != 0u) apply_check((unsigned)(rec[t] >> 32), (int)(unsigned)rec[t], old[t]);
          if (HAS_CNT && lane + 32u * t < cnt) count_one((unsigned)(rec[t] >> 32));
        }
      }
      __syncwarp();
      if (lane == 0) dsm::remote_arrive_relaxed(rempty0[j] + b * BAR_B);   // tell the sender of region j its buffer b is free
    }
  };

  // ---- main loop: phase = 128 elements per warp ------------------------------------------------------------------
  // the raw loads of the next phase stay in flight while the current one is worked on
  int4 l0, l1, av;                                             // current vector: 4 labels (two 16-byte halves for int64), 4 values
  auto load = [&](long long vv, int4& a0, int4& a1, int4& xv) {
    a0 = make_int4(0, 0, 0, 0); a1 = a0; xv = a0;
    if (vv < nvec) {
      if (sizeof(LT) == 8) { a0 = ldg_stream16((const int4*)p.labels + 2 * vv); a1 = ldg_stream16((const int4*)p.labels + 2 * vv + 1); }
      else a0 = ldg_stream16((const int4*)p.labels + vv);
      xv = ldg_stream16((const int4*)p.a + vv);
    }
  };
  load(v, l0, l1, av);
  const unsigned nph = (unsigned)nphase;
  for (unsigned ph = 0; ph < nph; ++ph) {
    int4 n0, n1, nv;
    load(v + grid_vecs, n0, n1, nv);
    const unsigned b = ph & 1u;
    if (ph >= 2) {                                             // buffer b is free once the peers applied phase ph - 2
      const unsigned par = ((ph >> 1) - 1u) & 1u;
#pragma unroll
      for (int j = 0; j < P; ++j) dsm::mbar_wait(empty0 + (j * 2 + b) * BAR_B, par);
    }
    unsigned dst[P], dbar[P], sent[P];
#pragma unroll
    for (int j = 0; j < P; ++j) { dst[j] = dst0[j] + b * buf_stride; dbar[j] = dbar0[j] + b * BAR_B; sent[j] = 0; }
    // labels as 32-bit words; for int64 labels all four high words must be zero or the vector takes the slow look
    unsigned glo[4], hv;
    if (sizeof(LT) == 8) { glo[0] = l0.x; glo[1] = l0.z; glo[2] = l1.x; glo[3] = l1.z; hv = (unsigned)(l0.y | l0.w | l1.y | l1.w); }
    else { glo[0] = l0.x; glo[1] = l0.y; glo[2] = l0.z; glo[3] = l0.w; hv = (unsigned)(l0.x | l0.y | l0.z | l0.w) >> 31; }
    const unsigned have = dsm::keep(v < nvec ? 1u : 0u);
    const unsigned covv = (have != 0u && hv == 0u) ? cov_c : 0u;
    const unsigned xb[4] = {(unsigned)av.x, (unsigned)av.y, (unsigned)av.z, (unsigned)av.w};
    // ---- step 1: classify the four elements (fixed-point value, owner, slot) ----
    int q4[4]; unsigned rel4[4]; unsigned fastm = 0u;          // bit k of fastm: element k goes into a table of the cluster
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned g = glo[k];
      const float x = __uint_as_float(xb[k]);
      const unsigned ab = xb[k] & 0x7fffffffu;
      bool fast = (ab - lo_bits < span_bits) && g < covv;
      int q = __float2int_rn(x * scale_f);                     // exact for in-window values
      if (!fast && have != 0u) {
        // ---- the rare cases, one element at a time.  Only a vector with a label outside [0, 2^32) reads its labels again
        //      (a dependent global load here stalled 40 % of the warp-phases when every out-of-window value took it) ----
        const long long g64 = hv == 0u ? (long long)g : LabelVec<LT>::one(p.labels, (v << 2) + k);
        if ((unsigned long long)g64 >= (unsigned long long)p.size) bad = 1;
        else if (!(NANSKIP && x != x)) {
          const bool slotted = (unsigned long long)g64 < (unsigned long long)cov_c;
          const bool inwin = ab - lo_bits < span_bits;
          if ((!slotted || !inwin) && ab != 0u) red_add_f64_hint(p.t.t0f + g64, (double)x, l2_evict_last_policy());   // exact f64 bypass
          if (HAS_CNT) {
            if (!slotted) global_count(g64);
            else if (!inwin) { fast = true; q = 0; }            // the count still travels to the owner
          }
          if (slotted && inwin) fast = true;                   // only this vector's OTHER labels were odd
        }
      }
      q4[k] = q;
      rel4[k] = (g - me) & (C - 1);                            // 0: mine, j + 1: peer j
      fastm |= fast ? (1u << k) : 0u;
    }
    // ---- step 2: records for the peers (no dependence on my own table) ----
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const bool go = ((fastm >> k) & 1u) && rel4[k] == (unsigned)(j + 1);
        const unsigned m = __ballot_sync(0xffffffffu, go);
        if (go) {
          const unsigned slot = 1u + sent[j] + __popc(m & lt_mask);
          if (slot < capm) dsm::st_async_rec(dst[j] + slot * 8u, (unsigned)q4[k], glo[k] >> LOGC, dbar[j]);
          else {                                               // buffer full: the slow exact way
            if (q4[k] != 0) atomicAdd(p.t.t0f + glo[k], (double)__uint_as_float(xb[k]));
            if (HAS_CNT) global_count((long long)glo[k]);
          }
        }
        sent[j] += __popc(m);
      }
    }
    // ---- step 3: my own elements: all ATOMS first, then the carry checks ----
    unsigned old4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      old4[k] = 0u;
      if (((fastm >> k) & 1u) && rel4[k] == 0u && (!HAS_CNT || q4[k] != 0)) old4[k] = apply_issue(glo[k] >> LOGC, q4[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (((fastm >> k) & 1u) && rel4[k] == 0u) {
        if (!HAS_CNT || q4[k] != 0) apply_check(glo[k] >> LOGC, q4[k], old4[k]);
        if (HAS_CNT) count_one(glo[k] >> LOGC);
      }
    }
    // announce the phase: header {count} + one arrival per peer
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const unsigned c = min(sent[j], capm - 1u);
        dsm::st_async_rec(dst[j], c, 0u, dbar[j]);
        dsm::remote_arrive_expect_tx(dbar[j], 8u * (c + 1u));
      }
    }
    if (ph >= 1) consume(ph - 1);
    v += grid_vecs;
    l0 = n0; l1 = n1; av = nv;
  }
  if (nph >= 1) consume(nph - 1);
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {            // ragged tail: straight to the global tables
    const long long i = (nvec << 2) + threadIdx.x;
    const long long g = LabelVec<LT>::one(p.labels, i);
    const float x = p.a[i];
    if ((unsigned long long)g >= (unsigned long long)p.size) bad = 1;
    else if (!(NANSKIP && x != x)) {
      if ((__float_as_uint(x) & 0x7fffffffu) != 0u) red_add_f64_hint(p.t.t0f + g, (double)x, l2_evict_last_policy());
      if (HAS_CNT) global_count(g);
    }
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.status[AGC_ST_REGIME] = 70 + MODE;
  dsm::cluster_sync();                // nobody leaves (or flushes) while a peer may still post into this CTA

  // ---- merge my slots into the global tables: ONE RED per touched group ----------------------------------------
  for (int li = threadIdx.x; li < cov; li += TEAM_THREADS) {
    const long long g = ((long long)li << LOGC) | me;
    const long long resid = (long long)table[li] - 0x80000000ll;
    if (resid != 0) atomicAdd(p.t.t0f + g, (double)resid * inv_scale);
    if (HAS_CNT) {
      const unsigned c = (cnt16[li >> 1] >> ((li & 1) << 4)) & 0xffffu;
      if (c) {
        if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + g), (unsigned long long)c);
        if (p.want_b1) p.t.b1[g] = 1;
      }
    }
  }
}

// shared memory one CTA needs
inline size_t team_smem_bytes(int C, int mode, int cov, int cap) {
  const int table_words = cov + (mode == TM_SUMCNT ? (cov + 1) / 2 : 0);
  const size_t inbox_off = (size_t)((table_words + 3) & ~3) * 4;
  return inbox_off + (size_t)(C - 1) * 2 * TEAM_WARPS * cap * 8 + (size_t)2 * (C - 1) * 2 * TEAM_WARPS * 8;
}
// record slots per (peer, buffer, warp): header + mean + 2.5 sigma of Binomial(128, 1/C), rounded up to a multiple of 4
inline int team_default_cap(int C) {
  const double mean = 128.0 / C, sd = sqrt(128.0 * (1.0 / C) * (1.0 - 1.0 / C));
  return ((int)(1 + mean + 2.5 * sd) + 3) & ~3;
}

template <int C, class LT, int MODE, bool NANSKIP>
inline int launch_team(const TeamParams& p, size_t smem, cudaStream_t st, int max_clusters = 0) {
  auto kern = k_team<C, LT, MODE, NANSKIP>;
  static std::atomic<unsigned long long> attr_done{0};
  if (!ensure_dyn_smem(kern, 227 * 1024 - 64, attr_done)) return -30;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(TEAM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3(C);
  // co-resident clusters of this kernel on this device (one CTA per SM: independent of the exact table size)
  static std::atomic<int> resident[AGC_MAX_DEVICES];
  const int dev = current_device();
  int nclusters = resident[dev].load(std::memory_order_relaxed);
  if (nclusters == 0) {
    if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); nclusters = -1; }
    resident[dev].store(nclusters, std::memory_order_relaxed);
  }
  if (nclusters < 1) return -31;
  if (max_clusters > 0 && nclusters > max_clusters) nclusters = max_clusters;
  cfg.gridDim = dim3(nclusters * C);
  AGC_COUNT_LAUNCH(1);
  if (cudaLaunchKernelEx(&cfg, kern, p) != cudaSuccess) return -32;
  return nclusters;
}

}  // namespace agc
