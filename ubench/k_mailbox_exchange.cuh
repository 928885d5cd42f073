// agc_mailbox.cuh -- tables that need the shared memory of K = 2 or 4 SMs: the SMs of a TEAM split the groups and
// hand each other the elements they do not own through small rings in L2.
//
// Why: a table of 10^5 one-word sums is 400 KB; one SM holds 226 KB.  The partial-table kernel (agc_fast.cuh) sends
// the uncovered 44 % of the elements to the global table as one RED packet each and is bound by the SM's RED issue
// rate (~0.6 packets/clk).  Here nothing leaves the chip as a scattered atomic: CTA `b` (one per SM, all co-resident:
// cooperative launch) owns the groups g with g % K == b % K, local index g / K.  Every CTA streams its share of the
// input as in k_fast and converts each value to its accumulator form (fixed-point word / order-preserving key).  An
// element it owns is added with a native shared-memory atomic; any other element becomes an 8-byte RECORD
// (local index, converted value) that the warp appends -- compacted with one ballot per element slot, so a warp's
// records of one slot land in consecutive words -- to a ring owned by (destination CTA, source CTA, warp number).
// Warp w of the destination drains the rings of warp w of its team mates once per iteration and applies the records
// to its own table.
//
// The rings need no fences and no tail pointers: a record is ONE 64-bit store whose bit 32 carries the lap parity of
// its slot (the memory starts zeroed, lap 0 writes parity 1, lap 1 parity 0, ...).  The consumer loads MB_J x 32 slots
// speculatively (ld.relaxed.gpu, L2), takes the contiguous prefix whose parity matches, and publishes its position in
// a `head` word; the producer prefetches `head` with its input loads and only waits when the ring could be full.  The
// head value is computed from the loaded records, so it cannot become visible before they were read.  A warp that has
// to wait for space keeps draining its own rings, which is what makes the exchange deadlock-free; a warp that has
// finished its input publishes how many records it wrote per ring and serves its own rings until it has consumed the
// totals its team mates published.  Rings are per warp, so there is no block barrier anywhere in the main loop.
//
// Formats: sums are the one-word carry-forwarding fixed point of FM_SUM1 (agc_fast.cuh) with ONE reference exponent
// per team (agreed through a team word before the main loop, so a record can carry the converted word); values
// outside the exact window go to the global f64 table from the CTA that read them.  mean / var add a u32 count word,
// max / min use the order-preserving s32 key, len a u32 count.
#pragma once
#include "agc_fast.cuh"

namespace agc {

constexpr int MB_THREADS = 1024;
constexpr int MB_J = 5;                          // 32-slot groups a warp loads per ring and iteration (production: 4 on average at K = 2)
constexpr unsigned MB_SPIN_LIMIT = 1u << 22;     // watchdog: a few seconds of polling, then status[1] is raised

struct MailboxParams {
  unsigned long long* ring;      // [ctas][K-1][32 warps][RING] records, zeroed before the launch
  unsigned* head;                // [ctas][K-1][32 warps] x 8 words: [0] consumer position, [4] producer's final count + 1
  unsigned* team;                // [teams] x 8 words: [0] max |a| bits, [1] arrivals
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// shared-memory atomics on a 32-bit shared-window address (spares the generic -> shared conversion at every site)
__device__ __forceinline__ unsigned atoms_add(unsigned addr, unsigned v) {
  unsigned old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void reds_add(unsigned addr, unsigned v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_max(unsigned addr, int v) { asm volatile("red.shared.max.s32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_min(unsigned addr, int v) { asm volatile("red.shared.min.s32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }

// MODE: FM_SUM1 (one-word sum), FM_SUMCNT (one-word sum + u32 count), FM_MAX, FM_MIN, FM_LEN
template <int MODE, class LT, bool LOAD_A, int K>
__global__ void __launch_bounds__(MB_THREADS, 1) k_mailbox(const FastParams p, const MailboxParams mb) {
  if (p.choose && *p.choose != p.want) return;
  extern __shared__ unsigned sw[];
  __shared__ unsigned s_maxbits;
  constexpr int LOGK = K == 2 ? 1 : 2;
  constexpr int NS = K - 1;                                   // streams in and out per warp
  constexpr int RLOG = K == 2 ? 9 : 8;                        // log2(records per ring)
  constexpr unsigned RING = 1u << RLOG, RMASK = RING - 1u;
  constexpr bool IS_SUM = MODE == FM_SUM1 || MODE == FM_SUMCNT;
  constexpr int WORDS = MODE == FM_SUMCNT ? 2 : 1;
  constexpr unsigned FULL = 0xffffffffu;
  const unsigned member = blockIdx.x & (K - 1);
  const unsigned team0 = blockIdx.x & ~(K - 1);
  const int L = (int)((p.size + K - 1) >> LOGK);              // local groups of a member (upper bound)
  const unsigned init0 = (MODE == FM_MAX || IS_SUM) ? 0x80000000u : (MODE == FM_MIN ? 0x7fffffffu : 0u);
  for (int i = threadIdx.x; i < WORDS * L; i += blockDim.x) sw[i] = i < L ? init0 : 0u;
  if (threadIdx.x == 0) s_maxbits = 0u;
  __syncthreads();

  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long nvec = p.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  const bool square = p.square != 0;
  unsigned flags = 0;                                         // 1: label out of range, 2: watchdog

  // ---- reference exponent of the TEAM: largest finite |a| of the members' first tiles -------------------
  unsigned lo_bits = 0, span_bits = 0; float scale_f = 1.f; double inv_scale = 1.0;
  if (IS_SUM) {
    unsigned m = 0;
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v < nvec) {
      int4 av = ldg_stream16((const int4*)p.a + v);
      unsigned b[4] = {(unsigned)av.x & 0x7fffffffu, (unsigned)av.y & 0x7fffffffu, (unsigned)av.z & 0x7fffffffu, (unsigned)av.w & 0x7fffffffu};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned bb = b[k];
        if (square) { float f = __uint_as_float(bb); bb = __float_as_uint(f * f); }
        if (bb < 0x7f800000u && bb > m) m = bb;
      }
    }
    m = __reduce_max_sync(FULL, m);
    if (lane == 0 && m) atomicMax(&s_maxbits, m);
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned* tw = mb.team + (team0 >> LOGK) * 8;
      atomicMax(tw, s_maxbits);
      __threadfence();
      atomicAdd(tw + 1, 1u);
      unsigned spins = 0;
      while (ld_relaxed_u32(tw + 1) < (unsigned)K && ++spins < MB_SPIN_LIMIT) __nanosleep(64);
      if (spins >= MB_SPIN_LIMIT) p.status[AGC_ST_PRECISION] = 0x7E57;
      __threadfence();
      s_maxbits = ld_relaxed_u32(tw);
    }
    __syncthreads();
    int eb = (int)(s_maxbits >> 23);
    if (eb == 0) eb = 127;
    eb += 1;
    if (eb < 30) eb = 30;                                     // keeps 2^(FX1_F - (eb-127)) a normal float; tinier data takes the bypass
    int lo_e = eb - FX1_WINDOW; if (lo_e < 1) lo_e = 1;
    lo_bits = (unsigned)lo_e << 23;
    span_bits = (unsigned)(eb - lo_e) << 23;
    scale_f = __uint_as_float((unsigned)(127 + FX1_F - (eb - 127)) << 23);      // q = x * 2^(FX1_F - (eb-127)): exact in the window
    inv_scale = __longlong_as_double((long long)(1023 - FX1_F + (eb - 127)) << 52);
  }
  const double hi_unit = inv_scale * 4294967296.0;            // what one carry out of the 32-bit word is worth

  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sw);
  // one converted element of a group this CTA owns (local index l)
  auto apply = [&](unsigned l, unsigned q) {
    const unsigned addr = sbase + l * 4u;
    if (MODE == FM_LEN) { reds_add(addr, 1u); return; }
    if (MODE == FM_MAX) { reds_max(addr, (int)q); return; }
    if (MODE == FM_MIN) { reds_min(addr, (int)q); return; }
    const unsigned old = atoms_add(addr, q);
    const int net = ((int)q >> 31) + (int)(old + q < old);    // carry (+1) / borrow (-1) out of the word
    if (net) atomicAdd(p.t.t0f + (((long long)l << LOGK) | member), net > 0 ? hi_unit : -hi_unit);
    if (MODE == FM_SUMCNT) reds_add(addr + (unsigned)L * 4u, 1u);
  };

  // ---- rings: in-stream j comes from member (member + 1 + j) % K; out-stream j goes to that member -------
  unsigned long long* in_ring[NS]; unsigned long long* out_ring[NS];
  unsigned* in_ctl[NS]; unsigned* out_ctl[NS];
  unsigned wpos[NS], hcache[NS], rpos[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const size_t si = (size_t)((blockIdx.x * NS + j) * 32 + w);
    const unsigned d = (member + 1 + j) & (K - 1);
    const unsigned jin = (member - d - 1) & (K - 1);
    const size_t so = (size_t)(((team0 + d) * NS + jin) * 32 + w);
    in_ring[j] = mb.ring + (si << RLOG); in_ctl[j] = mb.head + si * 8;
    out_ring[j] = mb.ring + (so << RLOG); out_ctl[j] = mb.head + so * 8;
    wpos[j] = 0; hcache[j] = 0; rpos[j] = 0;
  }

  auto drain = [&]() {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      unsigned long long rec[MB_J];
      const unsigned r0 = rpos[j];
#pragma unroll
      for (int i = 0; i < MB_J; ++i) rec[i] = ld_relaxed_u64(in_ring[j] + ((r0 + i * 32 + lane) & RMASK));
      bool open = true;
#pragma unroll
      for (int i = 0; i < MB_J; ++i) {
        if (open) {
          const unsigned pos = r0 + i * 32 + lane;
          const unsigned hi = (unsigned)(rec[i] >> 32);
          const unsigned vm = __ballot_sync(FULL, ((hi ^ (pos >> RLOG)) & 1u) != 0u);       // parity of the slot's lap
          const unsigned nv = vm == FULL ? 32u : (unsigned)__ffs(~vm) - 1u;
          if (lane < nv) apply(hi >> 1, (unsigned)rec[i]);
          rpos[j] += nv;
          open = nv == 32u;
        }
      }
      if (rpos[j] != r0 && lane == 0) st_relaxed_u32(in_ctl[j], rpos[j]);
    }
  };

  // make room for `cnt` more records on out-stream j; serves the own rings while it waits
  auto wait_space = [&](int j, unsigned cnt) {
    unsigned spins = 0;
    for (;;) {
      const unsigned h = ld_relaxed_u32(out_ctl[j]);          // same address in every lane: one request
      hcache[j] = h;
      if (wpos[j] + cnt - h <= RING || (flags & 2u)) break;
      drain();
      if (++spins > MB_SPIN_LIMIT) flags |= 2u;
      __nanosleep(32);
    }
  };
  auto put = [&](int j, unsigned pos, unsigned l, unsigned q) {
    const unsigned hi = (l << 1) | (((pos >> RLOG) & 1u) ^ 1u);
    st_relaxed_u64(out_ring[j] + (pos & RMASK), ((unsigned long long)hi << 32) | q);
  };

  // ---- main loop: warp-uniform trip count (ballots) ---------------------------------------------------
  const unsigned lt_mask = (1u << lane) - 1u;
  for (long long vb = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); vb < nvec; vb += 2 * stride) {
    long long g[2][4]; float x[2][4]; bool in[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long v = vb + u * stride + lane;
      in[u] = v < nvec;
#pragma unroll
      for (int k = 0; k < 4; ++k) { g[u][k] = -1; x[u][k] = 0.f; }
      if (in[u]) {
        LabelVec<LT>::load(p.labels, v, g[u]);
        if (LOAD_A) {
          int4 av = ldg_stream16((const int4*)p.a + v);
          x[u][0] = __int_as_float(av.x); x[u][1] = __int_as_float(av.y); x[u][2] = __int_as_float(av.z); x[u][3] = __int_as_float(av.w);
        }
      }
    }
    unsigned hnew[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) hnew[j] = ld_relaxed_u32(out_ctl[j]);
    drain();                                                  // the L2 round trip of the rings hides behind the DRAM loads
#pragma unroll
    for (int j = 0; j < NS; ++j) hcache[j] = hnew[j];

    // convert the 8 element slots; ballots per destination, then one space check per out-stream
    unsigned q[8];
    unsigned fm[NS][8];                                       // warp-uniform masks
    unsigned cnt[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) cnt[j] = 0;
    unsigned ownmask = 0;                                     // per-thread: slots this CTA owns
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const long long gg = g[e >> 2][e & 3];
      float xx = x[e >> 2][e & 3];
      bool ok = (unsigned long long)gg < (unsigned long long)p.size;
      if (!ok && in[e >> 2]) flags |= 1u;
      const bool nan = xx != xx;
      if (LOAD_A && nanskip && nan) ok = false;
      if (IS_SUM) {
        if (square) xx = xx * xx;
        const unsigned ab = __float_as_uint(xx) & 0x7fffffffu;
        if (ab - lo_bits < span_bits) q[e] = (unsigned)__float2int_rn(xx * scale_f);       // |q| < 2^29, exact
        else {
          q[e] = 0u;                                          // mean / var: the record still counts the element
          if (ok && ab != 0u) atomicAdd(p.t.t0f + gg, (double)xx);                         // out of window / inf / NaN: exact f64 RED
          if (MODE == FM_SUM1) ok = false;
        }
      } else if (MODE == FM_MAX) q[e] = (unsigned)(nan ? 0x7fffffff : key32_of(xx));        // NaN wins
      else if (MODE == FM_MIN) q[e] = (unsigned)(nan ? (int)0x80000000 : key32_of(xx));
      else q[e] = 0u;
      const unsigned d = (unsigned)gg & (K - 1);
      if (ok && d == member) ownmask |= 1u << e;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        fm[j][e] = __ballot_sync(FULL, ok && d == ((member + 1 + j) & (K - 1)));
        cnt[j] += __popc(fm[j][e]);
      }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) if (wpos[j] + cnt[j] - hcache[j] > RING) wait_space(j, cnt[j]);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const unsigned l = (unsigned)g[e >> 2][e & 3] >> LOGK;
      if ((ownmask >> e) & 1u) apply(l, q[e]);
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const unsigned m = fm[j][e];
        if ((m >> lane) & 1u) put(j, wpos[j] + __popc(m & lt_mask), l, q[e]);
        wpos[j] += __popc(m);
      }
    }
  }

  // ---- end of input: publish the totals, then serve the own rings until the team mates' totals are consumed ----
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NS; ++j) st_relaxed_u32(out_ctl[j] + 4, wpos[j] + 1u);
  }
  {
    unsigned spins = 0, pending = (1u << NS) - 1u;
    while (pending && !(flags & 2u)) {
      drain();
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        if ((pending >> j) & 1u) {
          const unsigned t = ld_relaxed_u32(in_ctl[j] + 4);
          if (t && rpos[j] == t - 1u) pending &= ~(1u << j);
        }
      }
      if (++spins > MB_SPIN_LIMIT) flags |= 2u;
    }
  }
  if (flags & 2u) p.status[AGC_ST_PRECISION] = 0x7E57;

  // ragged tail: straight to the global tables
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {
    const long long i = (nvec << 2) + threadIdx.x;
    const long long gg = LabelVec<LT>::one(p.labels, i);
    float xx = LOAD_A ? p.a[i] : 0.f;
    if ((unsigned long long)gg >= (unsigned long long)p.size) flags |= 1u;
    else if (!(LOAD_A && nanskip && xx != xx)) {
      if (MODE == FM_LEN) atomicAdd((unsigned long long*)(p.t.t1 + gg), 1ull);
      else if (MODE == FM_MAX || MODE == FM_MIN) {
        const long long k64 = xx != xx ? (MODE == FM_MAX ? AGC_EMPTY_MIN : AGC_EMPTY_MAX) : key_of((double)xx);
        if (MODE == FM_MAX) atomicMax(p.t.t0i + gg, k64); else atomicMin(p.t.t0i + gg, k64);
      } else {
        if (square) xx = xx * xx;
        if ((__float_as_uint(xx) & 0x7fffffffu) != 0u) atomicAdd(p.t.t0f + gg, (double)xx);
        if (MODE == FM_SUMCNT) { if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + gg), 1ull); if (p.want_b1) p.t.b1[gg] = 1; }
      }
    }
  }
  if (flags & 1u) p.status[AGC_ST_BADLABEL] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.status[AGC_ST_REGIME] = 30 + MODE;
  __syncthreads();

  // ---- merge: one RED per touched group of this member ------------------------------------------------
  const unsigned* w0 = sw; const unsigned* w1 = sw + L;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const long long gg = ((long long)l << LOGK) | member;
    if (gg >= p.size) continue;
    if (MODE == FM_LEN) {
      const unsigned c = w0[l];
      if (c) atomicAdd((unsigned long long*)(p.t.t1 + gg), (unsigned long long)c);
    } else if (MODE == FM_MAX || MODE == FM_MIN) {
      const int k = (int)w0[l];
      if (k != (int)init0) {
        long long k64;
        if (MODE == FM_MAX) k64 = k == 0x7fffffff ? AGC_EMPTY_MIN : key_of((double)unkey_f32(k));
        else k64 = k == (int)0x80000000 ? AGC_EMPTY_MAX : key_of((double)unkey_f32(k));
        if (MODE == FM_MAX) atomicMax(p.t.t0i + gg, k64); else atomicMin(p.t.t0i + gg, k64);
      }
    } else {
      const long long qq = (long long)w0[l] - 0x80000000ll;
      if (qq) atomicAdd(p.t.t0f + gg, (double)qq * inv_scale);
      if (MODE == FM_SUMCNT) {
        const unsigned c = w1[l];
        if (c) {
          if (p.want_cnt) atomicAdd((unsigned long long*)(p.t.t1 + gg), (unsigned long long)c);
          if (p.want_b1) p.t.b1[gg] = 1;
        }
      }
    }
  }
}

// workspace the rings need
inline int mailbox_ring_log2(int K) { return K == 2 ? 9 : 8; }
inline size_t mailbox_bytes(int sms, int K) {
  const size_t ctas = (size_t)(sms / K * K), subs = ctas * (K - 1) * 32;
  return (subs << mailbox_ring_log2(K)) * 8 + subs * 32 + (ctas / K) * 32;
}
constexpr size_t MAILBOX_MAX_BYTES = 48u << 20;

template <int MODE, class LT, bool LOAD_A, int K>
inline int launch_mailbox_k(FastParams p, void* scratch, cudaStream_t st, int sms) {
  auto kern = k_mailbox<MODE, LT, LOAD_A, K>;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64) != cudaSuccess) return -30;
    attr_done = true;
  }
  const int ctas = sms / K * K;
  const size_t subs = (size_t)ctas * (K - 1) * 32;
  MailboxParams mb;
  mb.ring = (unsigned long long*)scratch;
  mb.head = (unsigned*)((char*)scratch + (subs << mailbox_ring_log2(K)) * 8);
  mb.team = mb.head + subs * 8;
  if (cudaMemsetAsync(scratch, 0, mailbox_bytes(sms, K), st) != cudaSuccess) return -31;
  constexpr int WORDS = MODE == FM_SUMCNT ? 2 : 1;
  const size_t smem = (size_t)((p.size + K - 1) / K) * WORDS * 4;
  void* args[] = {(void*)&p, (void*)&mb};
  AGC_COUNT_LAUNCH(1);
  // cooperative: every CTA must be resident, its team mates wait for it
  if (cudaLaunchCooperativeKernel((const void*)kern, dim3(ctas), dim3(MB_THREADS), args, smem, st) != cudaSuccess) { cudaGetLastError(); return -32; }
  return 1;
}

// smallest team that holds the table; 0 if none does
inline int mailbox_team(long long size, int words) {
  const size_t cap = 227 * 1024 - 1024;
  for (int K = 2; K <= 4; K *= 2)
    if ((size_t)((size + K - 1) / K) * words * 4 <= cap) return K;
  return 0;
}

template <int MODE, class LT, bool LOAD_A>
inline int launch_mailbox(const FastParams& p, int K, void* scratch, cudaStream_t st, int sms) {
  return K == 2 ? launch_mailbox_k<MODE, LT, LOAD_A, 2>(p, scratch, st, sms) : launch_mailbox_k<MODE, LT, LOAD_A, 4>(p, scratch, st, sms);
}

}  // namespace agc
