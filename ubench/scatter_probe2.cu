// scatter_probe2.cu -- second micro-probe (see scatter_probe.cu): can everything be built on
// native 32-bit integer shared atomics?
//   P1  fixed-point f32 sum = 64-bit accumulator as two u32 words: ATOMS.ADD(lo) with return +
//       conditional ATOMS.ADD(hi) for carry/sign                                (G <= 25K / SM)
//   P2  table distributed over a thread-block cluster, remote updates with
//       red.shared::cluster (DSMEM atomics)                                      (G = 1e5 .. 4e5)
//   P3  8 elements per thread per step (more bytes in flight at 1 CTA / SM)
//   P4  mean = fixed-point sum + count  (3 words / group)
//   P5  sorted labels: thread-run combine + warp segmented reduce, one update per segment tail
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scatter_probe2 scatter_probe2.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <string>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
__global__ void gen_uniform(long long* idx, float* a, long long n, long long G, uint64_t seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint64_t h = mix64(i ^ seed);
    idx[i] = (long long)(h % (uint64_t)G);
    a[i] = (float)((mix64(h) >> 40) * (1.0 / 16777216.0));
  }
}
__global__ void gen_zipf(long long* idx, float* a, long long n, long long G, const double* cdf, uint64_t seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint64_t h = mix64(i ^ seed);
    double u = (h >> 11) * (1.0 / 9007199254740992.0);
    long long lo = 0, hi = G - 1;
    while (lo < hi) { long long mid = (lo + hi) >> 1; if (cdf[mid] < u) lo = mid + 1; else hi = mid; }
    idx[i] = (long long)(((unsigned __int128)lo * 2654435761ull + 12345ull) % (uint64_t)G);
    a[i] = (float)((mix64(h) >> 40) * (1.0 / 16777216.0));
  }
}
__global__ void ramp(long long* idx, long long n, long long G) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    idx[i] = (long long)(((unsigned __int128)i * G) / n);
}
__device__ __forceinline__ int4 ldg_stream(const int4* p) {
  int4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
    : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
#define LBL(lo, hi) (((long long)(unsigned)(lo)) | ((long long)(hi) << 32))

template <int U, class F>
__device__ __forceinline__ void streamU(const long long* __restrict__ idx, const float* __restrict__ a, long long n, F f) {
  long long nvec = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; v + (U - 1) * stride < nvec; v += U * stride) {
    int4 l0[U], l1[U], av[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      l0[u] = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * (v + u * stride));
      l1[u] = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * (v + u * stride) + 1);
      av[u] = ldg_stream(reinterpret_cast<const int4*>(a) + (v + u * stride));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      f(LBL(l0[u].x, l0[u].y), __int_as_float(av[u].x)); f(LBL(l0[u].z, l0[u].w), __int_as_float(av[u].y));
      f(LBL(l1[u].x, l1[u].y), __int_as_float(av[u].z)); f(LBL(l1[u].z, l1[u].w), __int_as_float(av[u].w));
    }
  }
  for (; v < nvec; v += stride) {
    int4 l0 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v), l1 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v + 1);
    int4 av = ldg_stream(reinterpret_cast<const int4*>(a) + v);
    f(LBL(l0.x, l0.y), __int_as_float(av.x)); f(LBL(l0.z, l0.w), __int_as_float(av.y));
    f(LBL(l1.x, l1.y), __int_as_float(av.z)); f(LBL(l1.z, l1.w), __int_as_float(av.w));
  }
}

// ---- P1/P4: fixed-point accumulate in shared: value*2^SH as int64 in (lo,hi) u32 words -------------
__device__ __forceinline__ void fx_add(unsigned* lo, unsigned* hi, long long q) {
  unsigned ql = (unsigned)q; int qh = (int)(q >> 32);
  unsigned old = atomicAdd(lo, ql);
  qh += (old + ql < old);                 // carry out of the low word
  if (qh != 0) atomicAdd(hi, (unsigned)qh);
}
template <int U, bool COUNT>
__global__ void k_fx(const long long* idx, const float* a, long long n, double* tab, unsigned long long* cnt, int G, float scale) {
  extern __shared__ unsigned sw[];
  unsigned* lo = sw; unsigned* hi = sw + G; unsigned* cc = sw + 2 * G;
  for (int i = threadIdx.x; i < (COUNT ? 3 : 2) * G; i += blockDim.x) sw[i] = 0u;
  __syncthreads();
  streamU<U>(idx, a, n, [&](long long g, float v) {
    long long q = __float2ll_rn(v * scale);
    fx_add(lo + g, hi + g, q);
    if (COUNT) atomicAdd(cc + g, 1u);
  });
  __syncthreads();
  double inv = 1.0 / (double)scale;
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    long long q = (long long)(((unsigned long long)hi[i] << 32) | lo[i]);
    if (q) atomicAdd(tab + i, (double)q * inv);
    if (COUNT && cc[i]) atomicAdd(cnt + i, (unsigned long long)cc[i]);
  }
}
// reference points re-measured with the same skeleton
template <int U>
__global__ void k_u32(const long long* idx, const float* a, long long n, double* tab, int G) {
  extern __shared__ unsigned sw[];
  for (int i = threadIdx.x; i < G; i += blockDim.x) sw[i] = 0u;
  __syncthreads();
  streamU<U>(idx, a, n, [&](long long g, float v) { atomicAdd(sw + g, (unsigned)__float_as_int(v) >> 20); });
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) if (sw[i]) atomicAdd(tab + i, (double)sw[i]);
}

// ---- P2: cluster-distributed table, DSMEM remote atomics ----------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned local, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void red_cluster_add(unsigned addr, unsigned v) {
  asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_cluster_max(unsigned addr, int v) {
  asm volatile("red.relaxed.cluster.shared::cluster.max.s32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
// table slot of group g lives in CTA (g % C) at word (g / C); C = cluster size (power of two)
template <int U, int LOGC, bool MAXOP>
__global__ void k_cluster(const long long* idx, const float* a, long long n, double* tab, int G) {
  extern __shared__ unsigned sw[];
  constexpr unsigned C = 1u << LOGC;
  cg::cluster_group cl = cg::this_cluster();
  unsigned rank = cl.block_rank();
  int slots = (G + C - 1) >> LOGC;
  for (int i = threadIdx.x; i < slots; i += blockDim.x) sw[i] = MAXOP ? 0x80000000u : 0u;
  cl.sync();
  unsigned base = smem_u32(sw);
  unsigned rbase[C];
#pragma unroll
  for (unsigned r = 0; r < C; ++r) rbase[r] = mapa(base, r);
  streamU<U>(idx, a, n, [&](long long g, float v) {
    unsigned owner = (unsigned)g & (C - 1), slot = (unsigned)g >> LOGC;
    unsigned rb = rbase[0];
#pragma unroll
    for (unsigned r = 1; r < C; ++r) rb = owner == r ? rbase[r] : rb;
    if (MAXOP) { int b = __float_as_int(v); b ^= (b >> 31) & 0x7fffffff; red_cluster_max(rb + slot * 4, b); }
    else red_cluster_add(rb + slot * 4, (unsigned)__float_as_int(v) >> 20);
  });
  cl.sync();
  for (int i = threadIdx.x; i < slots; i += blockDim.x) {
    unsigned w = sw[i]; long long g = ((long long)i << LOGC) | rank;
    if (g < G && w != (MAXOP ? 0x80000000u : 0u)) atomicAdd(tab + g, (double)w);
  }
}
template <int U, int LOGC, bool MAXOP>
static void launch_cluster(int sms, int threads, size_t smem, const long long* idx, const float* a, long long n, double* tab, int G) {
  auto kern = k_cluster<U, LOGC, MAXOP>;
  static bool once = false;
  if (!once) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); once = true; }
  cudaLaunchConfig_t cfg = {};
  int C = 1 << LOGC;
  cfg.gridDim = dim3((sms / C) * C); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kern, idx, a, n, tab, G));
}

// ---- P5: sorted labels -- thread combine + warp segmented scan, tails update the table -------------
__global__ void k_sorted(const long long* idx, const float* a, long long n, double* tab) {
  long long nvec = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  int lane = threadIdx.x & 31;
  for (long long v0 = blockIdx.x * (long long)blockDim.x; v0 < nvec; v0 += stride) {
    long long v = v0 + threadIdx.x;
    bool ok = v < nvec;
    long long g0 = -1, g1 = -1, g2 = -1, g3 = -1; double x0 = 0, x1 = 0, x2 = 0, x3 = 0;
    if (ok) {
      int4 l0 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v), l1 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v + 1);
      int4 av = ldg_stream(reinterpret_cast<const int4*>(a) + v);
      g0 = LBL(l0.x, l0.y); g1 = LBL(l0.z, l0.w); g2 = LBL(l1.x, l1.y); g3 = LBL(l1.z, l1.w);
      x0 = __int_as_float(av.x); x1 = __int_as_float(av.y); x2 = __int_as_float(av.z); x3 = __int_as_float(av.w);
    }
    // in-thread runs: everything that ends before the thread's last label is flushed directly
    if (g1 == g0) x1 += x0; else if (ok) atomicAdd(tab + g0, x0);
    if (g2 == g1) x2 += x1; else if (ok) atomicAdd(tab + g1, x1);
    if (g3 == g2) x3 += x2; else if (ok) atomicAdd(tab + g2, x2);
    // (g3, x3) = the thread's open run; segmented inclusive scan over the warp keyed by g3,
    // valid when the thread holds a single run (g0 == g3); otherwise it is a segment head.
    double s = x3; long long key = g3;   // monotone labels: equal keys o lanes apart => one run in between
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, s, o);
      long long kk = __shfl_up_sync(0xffffffffu, key, o);
      if (lane >= o && kk == key) s += t;
    }
    long long nxt = __shfl_down_sync(0xffffffffu, g0, 1);
    bool tail = (lane == 31) || (nxt != g3);
    if (ok && tail) atomicAdd(tab + g3, s);
  }
}

struct Timer { cudaEvent_t a, b; Timer() { cudaEventCreate(&a); cudaEventCreate(&b); } };
template <class L> static double time_ms(L launch, int warm = 2, int reps = 5) {
  Timer t; for (int i = 0; i < warm; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(t.a); launch(); cudaEventRecord(t.b); CK(cudaEventSynchronize(t.b));
    float ms; cudaEventElapsedTime(&ms, t.a, t.b); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  int lg = argc > 1 ? atoi(argv[1]) : 28;
  long long N = 1ll << lg;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("# device %s, %d SMs, N=2^%d\n", prop.name, sms, lg);
  long long* idx; float* a; double* tab; unsigned long long* cnt; double* cdf;
  CK(cudaMalloc(&idx, N * 8)); CK(cudaMalloc(&a, N * 4));
  const long long GMAX = 10000000;
  CK(cudaMalloc(&tab, GMAX * 8)); CK(cudaMalloc(&cnt, GMAX * 8)); CK(cudaMalloc(&cdf, GMAX * 8));
  const double bytes = (double)N * 12.0;
  const int SMEM_MAX = 227 * 1024;
  auto setsm = [&](const void* f) { CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX)); };
  setsm((const void*)k_fx<1, false>); setsm((const void*)k_fx<2, false>); setsm((const void*)k_fx<1, true>); setsm((const void*)k_fx<2, true>);
  setsm((const void*)k_u32<1>); setsm((const void*)k_u32<2>); setsm((const void*)k_u32<4>);
  auto report = [&](const char* dist, long long G, const std::string& name, double ms) {
    printf("%-8s G=%-9lld %-44s %8.3f ms  %7.1f Gelem/s  %7.1f GB/s\n", dist, G, name.c_str(), ms, N / ms * 1e-6, bytes / ms * 1e-6);
    fflush(stdout);
  };
  const float scale = 4294967296.0f;   // 2^32: values in [0,1) -> 32 fractional bits
  long long Gs[] = {100, 10000, 25000, 50000, 100000, 400000};
  for (int dist = 0; dist < 3; ++dist) {
    for (long long G : Gs) {
      const char* dn = dist == 0 ? "uniform" : dist == 1 ? "zipf1.1" : "sorted";
      if (dist == 2 && !(G == 100 || G == 100000)) continue;
      if (dist == 0) gen_uniform<<<sms * 8, 256>>>(idx, a, N, G, 100);
      else if (dist == 1) {
        std::vector<double> h(G); double s = 0;
        for (long long k = 0; k < G; ++k) { s += pow((double)(k + 1), -1.1); h[k] = s; }
        for (long long k = 0; k < G; ++k) h[k] /= s;
        CK(cudaMemcpy(cdf, h.data(), G * 8, cudaMemcpyHostToDevice));
        gen_zipf<<<sms * 8, 256>>>(idx, a, N, G, cdf, 100);
      } else { gen_uniform<<<sms * 8, 256>>>(idx, a, N, G, 100); ramp<<<sms * 8, 256>>>(idx, N, G); }
      CK(cudaDeviceSynchronize());
      if (dist == 2) {
        CK(cudaMemset(tab, 0, G * 8));
        report(dn, G, "P5 sorted: thread runs + warp seg-scan", time_ms([&] { k_sorted<<<sms * 8, 256>>>(idx, a, N, tab); }));
        continue;
      }
      // single-CTA tables
      for (int words = 1; words <= 3; ++words) {
        size_t sm = (size_t)G * 4 * words;
        if (sm > (size_t)SMEM_MAX - 512) continue;
        int ctas = sm > 110 * 1024 ? 1 : (sm > 70 * 1024 ? 2 : (sm > 50 * 1024 ? 3 : 4));
        for (int thr : {512, 1024}) {
          if (ctas * thr > 2048) continue;
          for (int U : {1, 2}) {
            std::string tag = " words=" + std::to_string(words) + " cta/sm=" + std::to_string(ctas) + " thr=" + std::to_string(thr) + " U=" + std::to_string(U);
            if (words == 1) {
              if (U == 1) report(dn, G, "P3 u32 add" + tag, time_ms([&] { k_u32<1><<<sms * ctas, thr, sm>>>(idx, a, N, tab, (int)G); }));
              else report(dn, G, "P3 u32 add" + tag, time_ms([&] { k_u32<2><<<sms * ctas, thr, sm>>>(idx, a, N, tab, (int)G); }));
            } else if (words == 2) {
              if (U == 1) report(dn, G, "P1 fixed-point sum" + tag, time_ms([&] { k_fx<1, false><<<sms * ctas, thr, sm>>>(idx, a, N, tab, cnt, (int)G, scale); }));
              else report(dn, G, "P1 fixed-point sum" + tag, time_ms([&] { k_fx<2, false><<<sms * ctas, thr, sm>>>(idx, a, N, tab, cnt, (int)G, scale); }));
            } else {
              if (U == 1) report(dn, G, "P4 fixed-point mean" + tag, time_ms([&] { k_fx<1, true><<<sms * ctas, thr, sm>>>(idx, a, N, tab, cnt, (int)G, scale); }));
              else report(dn, G, "P4 fixed-point mean" + tag, time_ms([&] { k_fx<2, true><<<sms * ctas, thr, sm>>>(idx, a, N, tab, cnt, (int)G, scale); }));
            }
          }
        }
      }
      // cluster-distributed tables
      if (G >= 50000) {
        for (int logc = 1; logc <= 3; ++logc) {
          int C = 1 << logc;
          size_t sm = (size_t)((G + C - 1) / C) * 4;
          if (sm > (size_t)SMEM_MAX - 512) continue;
          std::string tag = " C=" + std::to_string(C) + " thr=1024 U=2";
          if (logc == 1) {
            report(dn, G, "P2 cluster red.add.u32" + tag, time_ms([&] { launch_cluster<2, 1, false>(sms, 1024, sm, idx, a, N, tab, (int)G); }));
            report(dn, G, "P2 cluster red.max.s32" + tag, time_ms([&] { launch_cluster<2, 1, true>(sms, 1024, sm, idx, a, N, tab, (int)G); }));
          } else if (logc == 2) {
            report(dn, G, "P2 cluster red.add.u32" + tag, time_ms([&] { launch_cluster<2, 2, false>(sms, 1024, sm, idx, a, N, tab, (int)G); }));
            report(dn, G, "P2 cluster red.max.s32" + tag, time_ms([&] { launch_cluster<2, 2, true>(sms, 1024, sm, idx, a, N, tab, (int)G); }));
          } else {
            report(dn, G, "P2 cluster red.add.u32" + tag, time_ms([&] { launch_cluster<2, 3, false>(sms, 1024, sm, idx, a, N, tab, (int)G); }));
            report(dn, G, "P2 cluster red.max.s32" + tag, time_ms([&] { launch_cluster<2, 3, true>(sms, 1024, sm, idx, a, N, tab, (int)G); }));
          }
        }
      }
    }
  }
  return 0;
}
