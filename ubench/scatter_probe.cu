// scatter_probe.cu -- throwaway-but-kept micro-probe of the update primitives the
// aggregate kernels can be built from, measured on the real B200.  It is NOT part
// of the product path: it exists so DESIGN.md's regime choices cite measurements.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scatter_probe scatter_probe.cu
//   ./scatter_probe [log2N=28]
//
// Workload = Form 1 sum: labels int64[N] (uniform or Zipf(1.1) over G groups), values f32[N].
// Every variant reads the same 12 B/element; what differs is the per-element update.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <string>

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
// inverse-CDF Zipf: cdf[k] = P(rank <= k); rank -> group through a multiplicative bijection
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

__device__ __forceinline__ int4 ldg_stream(const int4* p) {
  int4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
    : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}

// Common element-streaming skeleton: each thread takes 4 consecutive elements per step
// (2x16 B of labels + 16 B of values), grid-stride.  F is the per-element update functor.
template <class F>
__device__ __forceinline__ void stream4(const long long* __restrict__ idx, const float* __restrict__ a,
                                        long long n, F f) {
  long long nvec = n >> 2;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += stride) {
    int4 l0 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v);
    int4 l1 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v + 1);
    int4 av = ldg_stream(reinterpret_cast<const int4*>(a) + v);
    f(((long long)(unsigned)l0.x) | ((long long)l0.y << 32), __int_as_float(av.x));
    f(((long long)(unsigned)l0.z) | ((long long)l0.w << 32), __int_as_float(av.y));
    f(((long long)(unsigned)l1.x) | ((long long)l1.y << 32), __int_as_float(av.z));
    f(((long long)(unsigned)l1.z) | ((long long)l1.w << 32), __int_as_float(av.w));
  }
}

// K0: streaming ceiling -- read everything, reduce to one number (no scatter at all)
__global__ void k_stream(const long long* idx, const float* a, long long n, double* out) {
  float s = 0.f; long long t = 0;
  stream4(idx, a, n, [&](long long g, float v) { s += v; t ^= g; });
  if (s == 123.456f && t == 77) out[0] = s;   // keep the loads alive
}
// K1: one global f64 RED per element (table in L2/HBM)
__global__ void k_red_f64(const long long* idx, const float* a, long long n, double* tab) {
  stream4(idx, a, n, [&](long long g, float v) { atomicAdd(tab + g, (double)v); });
}
// K1b: one global f32 RED per element
__global__ void k_red_f32(const long long* idx, const float* a, long long n, float* tab) {
  stream4(idx, a, n, [&](long long g, float v) { atomicAdd(tab + g, v); });
}
// K1c: global signed-int max RED (order-preserving key) -- the min/max path
__global__ void k_red_max32(const long long* idx, const float* a, long long n, int* tab) {
  stream4(idx, a, n, [&](long long g, float v) {
    int b = __float_as_int(v); b ^= (b >> 31) & 0x7fffffff; atomicMax(tab + g, b); });
}
// K1d: two REDs per element (mean = f64 sum + u64 count)
__global__ void k_red_mean(const long long* idx, const float* a, long long n, double* tab, unsigned long long* cnt) {
  stream4(idx, a, n, [&](long long g, float v) { atomicAdd(tab + g, (double)v); atomicAdd(cnt + g, 1ull); });
}
// K2: CTA-private shared table, f32 atomicAdd (CAS-spin in SASS), flushed with f64 REDs
__global__ void k_smem_f32(const long long* idx, const float* a, long long n, double* tab, int G) {
  extern __shared__ float st[];
  for (int i = threadIdx.x; i < G; i += blockDim.x) st[i] = 0.f;
  __syncthreads();
  stream4(idx, a, n, [&](long long g, float v) { atomicAdd(&st[g], v); });
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) if (st[i] != 0.f) atomicAdd(tab + i, (double)st[i]);
}
// K2b: CTA-private shared table, native 32-bit integer ATOMS (count / max-key style)
__global__ void k_smem_u32(const long long* idx, const float* a, long long n, double* tab, int G) {
  extern __shared__ unsigned su[];
  for (int i = threadIdx.x; i < G; i += blockDim.x) su[i] = 0u;
  __syncthreads();
  stream4(idx, a, n, [&](long long g, float v) { atomicAdd(&su[g], (unsigned)__float_as_int(v) >> 20); });
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) if (su[i]) atomicAdd(tab + i, (double)su[i]);
}
// K2c: native ATOMS.MAX.S32 on keys
__global__ void k_smem_max(const long long* idx, const float* a, long long n, int* tab, int G) {
  extern __shared__ int si[];
  for (int i = threadIdx.x; i < G; i += blockDim.x) si[i] = INT_MIN;
  __syncthreads();
  stream4(idx, a, n, [&](long long g, float v) {
    int b = __float_as_int(v); b ^= (b >> 31) & 0x7fffffff; atomicMax(&si[g], b); });
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) if (si[i] != INT_MIN) atomicMax(tab + i, si[i]);
}
// K3: hybrid -- groups < Gs live in the CTA's shared table (native int ATOMS stands in for the
// update), the rest go to L2 with f64 RED.  Tests whether the two pipes add up.
__global__ void k_hybrid(const long long* idx, const float* a, long long n, double* tab, int Gs) {
  extern __shared__ unsigned su[];
  for (int i = threadIdx.x; i < Gs; i += blockDim.x) su[i] = 0u;
  __syncthreads();
  stream4(idx, a, n, [&](long long g, float v) {
    if (g < Gs) atomicAdd(&su[g], (unsigned)__float_as_int(v) >> 20);
    else atomicAdd(tab + g, (double)v); });
  __syncthreads();
  for (int i = threadIdx.x; i < Gs; i += blockDim.x) if (su[i]) atomicAdd(tab + i, (double)su[i]);
}
// K4: lane-private columns: slot[(warp, g, lane)] += v with plain LDS/FADD/STS (no atomics, no
// bank conflicts: bank == lane).  Needs warps*G*32*4 B of shared memory.
__global__ void k_lanepriv(const long long* idx, const float* a, long long n, double* tab, int G) {
  extern __shared__ float sl[];
  int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < nw * G * 32; i += blockDim.x) sl[i] = 0.f;
  __syncthreads();
  float* mine = sl + (size_t)w * G * 32 + lane;
  stream4(idx, a, n, [&](long long g, float v) { mine[(int)g * 32] += v; });
  __syncthreads();
  for (int g = w; g < G; g += nw) {             // one warp per group: reduce 32 lanes x nw warps
    double s = 0.0;
    for (int ww = 0; ww < nw; ++ww) s += (double)sl[((size_t)ww * G + g) * 32 + lane];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) atomicAdd(tab + g, s);
  }
}
// K5: plain (racy, wrong) LDS+FADD+STS on one CTA table: upper bound for any claim-style
// scheme that manages to make the update race-free without ATOMS.
__global__ void k_smem_plain(const long long* idx, const float* a, long long n, double* tab, int G) {
  extern __shared__ float st[];
  for (int i = threadIdx.x; i < G; i += blockDim.x) st[i] = 0.f;
  __syncthreads();
  stream4(idx, a, n, [&](long long g, float v) { st[g] += v; });
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) if (st[i] != 0.f) atomicAdd(tab + i, (double)st[i]);
}
// K6: MATCH.ANY cost alone (no memory update)
__global__ void k_match(const long long* idx, const float* a, long long n, double* out) {
  unsigned acc = 0;
  stream4(idx, a, n, [&](long long g, float v) { acc += __popc(__match_any_sync(0xffffffffu, (int)g)); });
  if (acc == 0xdeadbeef) out[0] = acc;
}
// K7: warp-aggregated RED: match peers, leader issues one f64 RED (value is not the true
// peer sum -- throughput probe only)
__global__ void k_match_red(const long long* idx, const float* a, long long n, double* tab) {
  int lane = threadIdx.x & 31;
  stream4(idx, a, n, [&](long long g, float v) {
    unsigned peers = __match_any_sync(0xffffffffu, (int)g);
    if ((__ffs(peers) - 1) == lane) atomicAdd(tab + g, (double)v * __popc(peers)); });
}
// K8: adjacent-run combine inside the thread's 4 elements + whole-warp-same-label shortcut,
// then f64 RED.  This is the "sorted / long runs" fast path riding on the generic kernel.
__global__ void k_runs_red(const long long* idx, const float* a, long long n, double* tab) {
  long long nvec = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += stride) {
    int4 l0 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v);
    int4 l1 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v + 1);
    int4 av = ldg_stream(reinterpret_cast<const int4*>(a) + v);
    long long g0 = ((long long)(unsigned)l0.x) | ((long long)l0.y << 32), g1 = ((long long)(unsigned)l0.z) | ((long long)l0.w << 32);
    long long g2 = ((long long)(unsigned)l1.x) | ((long long)l1.y << 32), g3 = ((long long)(unsigned)l1.z) | ((long long)l1.w << 32);
    double x0 = __int_as_float(av.x), x1 = __int_as_float(av.y), x2 = __int_as_float(av.z), x3 = __int_as_float(av.w);
    bool same = (g0 == g1) & (g1 == g2) & (g2 == g3);
    long long gl = __shfl_sync(0xffffffffu, g0, 0);
    if (__all_sync(0xffffffffu, same && g0 == gl)) {
      double s = (x0 + x1) + (x2 + x3);
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(tab + g0, s);
    } else {
      if (g1 == g0) { x1 += x0; } else atomicAdd(tab + g0, x0);
      if (g2 == g1) { x2 += x1; } else atomicAdd(tab + g1, x1);
      if (g3 == g2) { x3 += x2; } else atomicAdd(tab + g2, x2);
      atomicAdd(tab + g3, x3);
    }
  }
}

__global__ void ramp(long long* idx, long long n, long long G) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    idx[i] = (long long)(((unsigned __int128)i * G) / n);
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
  long long* idx; float* a; double* tab; unsigned long long* cnt;
  CK(cudaMalloc(&idx, N * 8)); CK(cudaMalloc(&a, N * 4));
  const long long GMAX = 10000000;
  CK(cudaMalloc(&tab, GMAX * 8)); CK(cudaMalloc(&cnt, GMAX * 8));
  double* cdf; CK(cudaMalloc(&cdf, GMAX * 8));
  const double bytes = (double)N * 12.0;
  const int SMEM_MAX = 227 * 1024;
  auto setsm = [&](const void* f) { CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX)); };
  setsm((const void*)k_smem_f32); setsm((const void*)k_smem_u32); setsm((const void*)k_smem_max);
  setsm((const void*)k_hybrid); setsm((const void*)k_lanepriv); setsm((const void*)k_smem_plain);

  auto report = [&](const char* dist, long long G, const char* name, double ms) {
    printf("%-8s G=%-9lld %-28s %8.3f ms  %7.1f Gelem/s  %7.1f GB/s\n", dist, G, name, ms, N / ms * 1e-6, bytes / ms * 1e-6);
    fflush(stdout);
  };
  long long Gs[] = {100, 1000, 10000, 50000, 100000, 1000000, 10000000};
  for (int dist = 0; dist < 3; ++dist) {
    for (long long G : Gs) {
      const char* dn = dist == 0 ? "uniform" : dist == 1 ? "zipf1.1" : "sorted";
      if (dist == 2 && !(G == 100 || G == 100000 || G == 10000000)) continue;
      if (dist == 0) gen_uniform<<<sms * 8, 256>>>(idx, a, N, G, 100);
      else if (dist == 1) {
        std::vector<double> h(G); double s = 0;
        for (long long k = 0; k < G; ++k) { s += pow((double)(k + 1), -1.1); h[k] = s; }
        for (long long k = 0; k < G; ++k) h[k] /= s;
        CK(cudaMemcpy(cdf, h.data(), G * 8, cudaMemcpyHostToDevice));
        gen_zipf<<<sms * 8, 256>>>(idx, a, N, G, cdf, 100);
      } else {
        // sorted labels: element i belongs to group floor(i*G/N)
        gen_uniform<<<sms * 8, 256>>>(idx, a, N, G, 100);   // values
        ramp<<<sms * 8, 256>>>(idx, N, G);
      }
      CK(cudaDeviceSynchronize());
      int gridA = sms * 8, blk = 256;
      if (dist == 0 && G == 100)
        report(dn, G, "K0 stream-only", time_ms([&] { k_stream<<<gridA, blk>>>(idx, a, N, tab); }));
      report(dn, G, "K1 red.f64", time_ms([&] { k_red_f64<<<gridA, blk>>>(idx, a, N, tab); }));
      report(dn, G, "K8 runs+red.f64", time_ms([&] { k_runs_red<<<gridA, blk>>>(idx, a, N, tab); }));
      if (dist == 2) continue;
      report(dn, G, "K1b red.f32", time_ms([&] { k_red_f32<<<gridA, blk>>>(idx, a, N, (float*)tab); }));
      report(dn, G, "K1c red.max.s32", time_ms([&] { k_red_max32<<<gridA, blk>>>(idx, a, N, (int*)tab); }));
      report(dn, G, "K1d red.f64+red.u64 (mean)", time_ms([&] { k_red_mean<<<gridA, blk>>>(idx, a, N, tab, cnt); }));
      report(dn, G, "K6 match.any only", time_ms([&] { k_match<<<gridA, blk>>>(idx, a, N, tab); }));
      report(dn, G, "K7 match.any+leader red", time_ms([&] { k_match_red<<<gridA, blk>>>(idx, a, N, tab); }));
      if (G * 4 <= SMEM_MAX - 1024) {
        size_t sm = (size_t)G * 4;
        int ctas = sm > 110 * 1024 ? 1 : (sm > 70 * 1024 ? 2 : 4);
        int b2 = ctas == 1 ? 1024 : 512;
        report(dn, G, "K2 smem atomicAdd.f32(CAS)", time_ms([&] { k_smem_f32<<<sms * ctas, b2, sm>>>(idx, a, N, tab, (int)G); }));
        report(dn, G, "K2b smem ATOMS.ADD.u32", time_ms([&] { k_smem_u32<<<sms * ctas, b2, sm>>>(idx, a, N, tab, (int)G); }));
        report(dn, G, "K2c smem ATOMS.MAX.s32", time_ms([&] { k_smem_max<<<sms * ctas, b2, sm>>>(idx, a, N, (int*)tab, (int)G); }));
        report(dn, G, "K5 smem plain RMW (racy)", time_ms([&] { k_smem_plain<<<sms * ctas, b2, sm>>>(idx, a, N, tab, (int)G); }));
      }
      if (G >= 100000) {
        for (int Gs2 : {16384, 32768, 49152}) {
          std::string nm = "K3 hybrid smem<" + std::to_string(Gs2) + " else red.f64";
          report(dn, G, nm.c_str(), time_ms([&] { k_hybrid<<<sms, 1024, (size_t)Gs2 * 4>>>(idx, a, N, tab, Gs2); }));
        }
      }
      if (G <= 1000) {
        for (int nw : {4, 8, 16}) {
          size_t sm = (size_t)nw * G * 32 * 4;
          if (sm > (size_t)SMEM_MAX) continue;
          int ctas = (int)((size_t)SMEM_MAX / (sm + 1024)); if (ctas < 1) ctas = 1; if (ctas * nw > 64) ctas = 64 / nw;
          std::string nm = "K4 lane-private w=" + std::to_string(nw) + " cta/sm=" + std::to_string(ctas);
          report(dn, G, nm.c_str(), time_ms([&] { k_lanepriv<<<sms * ctas, nw * 32, sm>>>(idx, a, N, tab, (int)G); }));
        }
      }
    }
  }
  return 0;
}
