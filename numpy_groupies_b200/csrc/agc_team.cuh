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
  int square;           // sumofsquares
  int want_cnt;         // TM_SUMCNT: merge counts into T1
  int want_b1;          // TM_SUMCNT: write touched flags
  int cov;              // table slots per CTA (group g -> CTA g % C, slot g / C; slots >= cov go to the global tables)
  int cap;              // 8-byte record slots per (peer, buffer, warp), slot 0 is the header {count}
  const int* choose;    // device word written by k_skew_probe; run iff *choose == want (NULL: always)
  int want;
};

namespace dsm {
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

// XPORT 0: st.async records + relaxed arrive.expect_tx (no fence anywhere in the loop)
// XPORT 1: plain st.shared::cluster records + release arrive (a fence per warp and phase) -- the fallback transport
template <int C, class LT, int MODE, int XPORT>
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
  const unsigned lt_mask = (1u << lane) - 1u;
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
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  const bool square = p.square != 0;
  bool bad = false;

  // ---- one reference exponent per cluster (records carry fixed-point integers) ----------------------------------
  {
    unsigned m = 0;
    if (v < nvec) {
      const int4 av = ldg_stream16((const int4*)p.a + v);
      const unsigned b[4] = {(unsigned)av.x & 0x7fffffffu, (unsigned)av.y & 0x7fffffffu, (unsigned)av.z & 0x7fffffffu, (unsigned)av.w & 0x7fffffffu};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned bb = b[k];
        if (square) { const float f = __uint_as_float(bb); bb = __float_as_uint(f * f); }
        if (bb < 0x7f800000u && bb > m) m = bb;
      }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0 && m) atomicMax(&s_maxbits, m);
    __syncthreads();
    if (threadIdx.x < C) dsm::st_remote_u32(dsm::mapa(dsm::smem_u32(&s_peer_max[me]), threadIdx.x), s_maxbits);
  }
  dsm::cluster_sync();                // mbarrier inits and the exponent words are visible cluster-wide behind this
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
  const double hi_unit = inv_scale * 4294967296.0;
  const unsigned long long pol = l2_evict_last_policy();
  const unsigned cnt_sbase = dsm::smem_u32(cnt16);

  // ---- per-peer addresses: peer j = CTA (me + j + 1) % C receives into ITS region j; I receive region j from CTA (me - j - 1) % C,
  //      which is peer P - 1 - j.  The shared::cluster window of a CTA is linear: remote address = peer base + local offset.
  unsigned rbase[P];
  const unsigned sw_s = dsm::smem_u32(sw);
#pragma unroll
  for (int j = 0; j < P; ++j) rbase[j] = dsm::mapa(sw_s, (me + j + 1) & (C - 1)) - sw_s;
  constexpr unsigned BAR_B = TEAM_WARPS * 8;                   // bytes from a buffer-0 barrier to its buffer-1 twin
  const unsigned buf_stride = (unsigned)(TEAM_WARPS * cap * 8);  // bytes from buffer 0 to buffer 1 of a region
  const unsigned in0 = dsm::smem_u32(inbox) + (unsigned)warp * cap * 8u;      // my region 0, buffer 0 (region j: + j * 2 * buf_stride)
  const unsigned full0 = dsm::smem_u32(bars) + (unsigned)warp * 8u;           // full[0][0][warp] (region j: + j * 2 * BAR_B)
  const unsigned empty0 = full0 + P * 2 * BAR_B;                             // empty[0][0][warp]

  // ---- element sinks ---------------------------------------------------------------------------------------
  auto repair = [&](unsigned wi, int sh, long long g) {        // k_len16's read-check-repair: take 0x8000 out, forward it
    unsigned o = *(volatile unsigned*)(cnt16 + wi);
    while (((o >> sh) & 0xffffu) >= 0x8000u) {
      const unsigned seen = atomicCAS(cnt16 + wi, o, o - (0x8000u << sh));
      if (seen == o) { red_add_u64_hint((unsigned long long*)(p.t.t1 + g), 0x8000ull, pol); break; }
      o = seen;
    }
  };
  auto apply = [&](unsigned li, int q) {                       // a record (or an own element) into MY table
    if (!HAS_CNT || q != 0) {
      const unsigned old = atomicAdd(table + li, (unsigned)q);
      const int net = (q >> 31) + (int)(old + (unsigned)q < old);
      if (net) red_add_f64_hint(p.t.t0f + (((long long)li << LOGC) | me), net > 0 ? hi_unit : -hi_unit, pol);
    }
    if (HAS_CNT) {
      const int sh = (int)(li & 1u) << 4; const unsigned wi = li >> 1;
      if (((*(volatile unsigned*)(cnt16 + wi) >> sh) & 0xffffu) >= 0x8000u) repair(wi, sh, ((long long)li << LOGC) | me);
      reds_add_u32(cnt_sbase + wi * 4, 1u << sh);
    }
  };
  auto global_add = [&](long long g, float x, bool with_count) {   // out-of-window values, slot-less groups, overflow records
    if ((__float_as_uint(x) & 0x7fffffffu) != 0u) red_add_f64_hint(p.t.t0f + g, (double)x, pol);
    if (HAS_CNT && with_count) {
      if (p.want_cnt) red_add_u64_hint((unsigned long long*)(p.t.t1 + g), 1ull, pol);
      if (p.want_b1) p.t.b1[g] = 1;
    }
  };
  auto consume = [&](long long ph) {                           // apply what the peers sent me in phase ph
    const unsigned b = (unsigned)ph & 1u, par = (unsigned)(ph >> 1) & 1u;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      dsm::mbar_wait(full0 + (j * 2 + b) * BAR_B, par);
      const unsigned long long* reg = inbox + ((size_t)(j * 2 + b) * TEAM_WARPS + warp) * cap;
      const unsigned cnt = (unsigned)reg[0];
      for (unsigned i = lane; i < cnt; i += 32) {
        const unsigned long long rec = reg[1 + i];
        apply((unsigned)(rec >> 32), (int)(unsigned)rec);
      }
      __syncwarp();
      // region j came from CTA (me - j - 1) % C = peer P - 1 - j; its "empty" barrier for ME is its index j
      if (lane == 0) dsm::remote_arrive_relaxed(rbase[P - 1 - j] + empty0 + (j * 2 + b) * BAR_B);
    }
  };

  // ---- main loop: phase = 128 elements per warp ------------------------------------------------------------------
  // raw loads of the next phase stay in flight while the current one is worked on
  long long g_raw[4]; float x_cur[4]; unsigned g_cur[4]; bool have_cur;
  auto load = [&](long long vv, long long (&g)[4], float (&x)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { g[k] = 0; x[k] = 0.f; }
    if (vv < nvec) {
      LabelVec<LT>::load(p.labels, vv, g);
      const int4 av = ldg_stream16((const int4*)p.a + vv);
      x[0] = __int_as_float(av.x); x[1] = __int_as_float(av.y); x[2] = __int_as_float(av.z); x[3] = __int_as_float(av.w);
    }
  };
  // labels -> 32 bits (size < 2^31): 0xffffffff marks a label outside [0, size)
  auto narrow = [&](const long long (&g)[4], unsigned (&o)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = (unsigned long long)g[k] < (unsigned long long)p.size ? (unsigned)g[k] : 0xffffffffu;
  };
  load(v, g_raw, x_cur);
  narrow(g_raw, g_cur);
  have_cur = v < nvec;
  for (long long ph = 0; ph < nphase; ++ph) {
    long long g_nxt[4]; float x_nxt[4];
    load(v + grid_vecs, g_nxt, x_nxt);
    const unsigned b = (unsigned)ph & 1u;
    if (ph >= 2) {                                             // buffer b is free once the peers applied phase ph - 2
      const unsigned par = (unsigned)((ph >> 1) - 1) & 1u;
#pragma unroll
      for (int j = 0; j < P; ++j) dsm::mbar_wait(empty0 + (j * 2 + b) * BAR_B, par);
    }
    unsigned sent[P];
#pragma unroll
    for (int j = 0; j < P; ++j) sent[j] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned g = g_cur[k];
      float x = x_cur[k];
      bool act = have_cur && g != 0xffffffffu;
      if (have_cur && !act) bad = true;
      if (nanskip && x != x) act = false;
      if (square) x = x * x;
      const unsigned ab = __float_as_uint(x) & 0x7fffffffu;
      const bool inwin = ab - lo_bits < span_bits;
      const unsigned li = g >> LOGC;
      if (act && li >= (unsigned)cov) { global_add((long long)g, x, true); act = false; }   // no slot for this group
      int q = 0;
      if (inwin) q = __float2int_rn(x * scale_f);                                            // exact
      else if (act) {
        if (ab != 0u) red_add_f64_hint(p.t.t0f + g, (double)x, pol);                          // exact bypass
        if (!HAS_CNT) act = false;                                                            // counts still travel (q = 0)
      }
      const unsigned rel = (g - me) & (C - 1);                                                // 0: mine, j + 1: peer j
      if (act && rel == 0u) apply(li, q);
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const bool go = act && rel == (unsigned)(j + 1);
        const unsigned m = __ballot_sync(0xffffffffu, go);
        if (go) {
          const unsigned slot = 1u + sent[j] + __popc(m & lt_mask);
          if (slot < (unsigned)cap) {
            const unsigned ra = rbase[j] + in0 + (j * 2 + b) * buf_stride + slot * 8u;
            if (XPORT == 0) dsm::st_async_rec(ra, (unsigned)q, li, rbase[j] + full0 + (j * 2 + b) * BAR_B);
            else dsm::st_remote_rec(ra, (unsigned)q, li);
          } else {                                                                            // buffer full: the slow exact way
            if (inwin) red_add_f64_hint(p.t.t0f + g, (double)x, pol);
            if (HAS_CNT) { if (p.want_cnt) red_add_u64_hint((unsigned long long*)(p.t.t1 + g), 1ull, pol); if (p.want_b1) p.t.b1[g] = 1; }
          }
        }
        sent[j] += __popc(m);
      }
    }
    // announce the phase: header {count} + one arrival per peer
    if (XPORT != 0) __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const unsigned c = min(sent[j], (unsigned)cap - 1u);
        const unsigned ra = rbase[j] + in0 + (j * 2 + b) * buf_stride;
        const unsigned rb = rbase[j] + full0 + (j * 2 + b) * BAR_B;
        if (XPORT == 0) {
          dsm::st_async_rec(ra, c, 0u, rb);
          dsm::remote_arrive_expect_tx(rb, 8u * (c + 1u));
        } else {
          dsm::st_remote_rec(ra, c, 0u);
          dsm::remote_arrive_release(rb);
        }
      }
    }
    if (ph >= 1) consume(ph - 1);
    v += grid_vecs;
    narrow(g_nxt, g_cur);
#pragma unroll
    for (int k = 0; k < 4; ++k) x_cur[k] = x_nxt[k];
    have_cur = v < nvec;
  }
  if (nphase >= 1) consume(nphase - 1);
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {            // ragged tail: straight to the global tables
    const long long i = (nvec << 2) + threadIdx.x;
    const long long g = LabelVec<LT>::one(p.labels, i);
    float x = p.a[i];
    if ((unsigned long long)g >= (unsigned long long)p.size) bad = true;
    else if (!(nanskip && x != x)) { if (square) x = x * x; global_add(g, x, true); }
  }
  if (bad) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.status[AGC_ST_REGIME] = 70 + MODE;
  dsm::cluster_sync();                // nobody leaves (or flushes) while a peer may still post into this CTA

  // ---- merge my slots into the global tables: ONE RED per touched group ----------------------------------------
  for (int li = threadIdx.x; li < cov; li += TEAM_THREADS) {
    const long long g = ((long long)li << LOGC) | me;
    if (g >= p.size) break;
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

template <int C, class LT, int MODE, int XPORT>
inline int launch_team(const TeamParams& p, size_t smem, cudaStream_t st, int max_clusters = 0) {
  auto kern = k_team<C, LT, MODE, XPORT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64) != cudaSuccess) return -30;
  if (C > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return -30;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(TEAM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3(C);
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); return -31; }
  if (max_clusters > 0 && nclusters > max_clusters) nclusters = max_clusters;
  cfg.gridDim = dim3(nclusters * C);
  AGC_COUNT_LAUNCH(1);
  if (cudaLaunchKernelEx(&cfg, kern, p) != cudaSuccess) return -32;
  return nclusters;
}

}  // namespace agc
