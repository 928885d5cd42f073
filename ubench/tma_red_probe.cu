// tma_red_probe.cu -- round-2 micro-probe (not part of the product path): can the TMA unit take scattered reductions off the
// LSU?  Every element is "uncovered" (no shared-memory table at all) and reaches the global f64 table either as a plain
// `red.global.add.f64` (mode 0), as a 16-byte `cp.reduce.async.bulk.global.shared::cta.add.f64` whose two-double record
// {x, 0} / {0, x} was staged in a per-thread shared-memory slot (mode 1), or half / half by lane parity (mode 2: do the
// two paths add up?).  The VERDICT's direction (b) -- flush staged records with bulk reductions -- needs the per-record
// rate of that instruction; nothing in the guides states it.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tma_red_probe tma_red_probe.cu && ./tma_red_probe [log2N=26] [G=100000]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>

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
__device__ __forceinline__ int4 ldg_stream(const int4* p) {
  int4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
    : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
__device__ __forceinline__ void red_f64(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

// DEPTH 16-byte slots per thread; a slot is reused once the bulk reduction that read it has done so (wait_group.read)
template <int MODE, int DEPTH>
__global__ void __launch_bounds__(1024, 1) k_tma_red(const long long* __restrict__ idx, const float* __restrict__ a, long long n, double* tab) {
  extern __shared__ __align__(16) double2 slots[];
  double2* mine = slots + threadIdx.x * DEPTH;
  const unsigned mine_s = (unsigned)__cvta_generic_to_shared(mine);
  const long long nvec = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  int fill = 0;
  const bool via_tma = MODE == 1 || (MODE == 2 && (threadIdx.x & 1));
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int4 l0 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v);
    const int4 l1 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v + 1);
    const int4 av = ldg_stream(reinterpret_cast<const int4*>(a) + v);
    const int g4[4] = {l0.x, l0.z, l1.x, l1.z};
    const int x4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = g4[j];
      const double x = (double)__int_as_float(x4[j]);
      if (!via_tma) { red_f64(tab + g, x); continue; }
      if (fill >= DEPTH) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(DEPTH - 1) : "memory");
      const int s = fill % DEPTH;
      mine[s] = (g & 1) ? make_double2(0.0, x) : make_double2(x, 0.0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 16;"
                   :: "l"(tab + (g & ~1)), "r"(mine_s + 16u * s) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      ++fill;
    }
  }
  if (via_tma) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


// What one uncovered element of `mean` costs the receiving side: WIDE 0 = two scalar REDs (f64 sum + u64 count, what the fused
// table format pays today), 1 = one scalar f32 RED, 2 = one red.global.add.v2.f32 {x, 1}, 4 = one red.global.add.v4.f32
// {x, 1, 0, 0} (one 16-byte slot per group: is a vector RED one operation of the L2 atomic unit or four?)
template <int WIDE>
__global__ void __launch_bounds__(1024, 1) k_wide_red(const long long* __restrict__ idx, const float* __restrict__ a, long long n, void* tabv) {
  const long long nvec = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int4 l0 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v);
    const int4 l1 = ldg_stream(reinterpret_cast<const int4*>(idx) + 2 * v + 1);
    const int4 av = ldg_stream(reinterpret_cast<const int4*>(a) + v);
    const int g4[4] = {l0.x, l0.z, l1.x, l1.z};
    const int x4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long g = g4[j];
      const float x = __int_as_float(x4[j]);
      if (WIDE == 0) {
        red_f64(reinterpret_cast<double*>(tabv) + 2 * g, (double)x);
        asm volatile("red.global.add.u64 [%0], %1;" :: "l"(reinterpret_cast<unsigned long long*>(tabv) + 2 * g + 1), "l"(1ull) : "memory");
      } else if (WIDE == 1) {
        asm volatile("red.global.add.f32 [%0], %1;" :: "l"(reinterpret_cast<float*>(tabv) + g), "f"(x) : "memory");
      } else if (WIDE == 2) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(reinterpret_cast<float*>(tabv) + 2 * g), "f"(x), "f"(1.0f) : "memory");
      } else {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(reinterpret_cast<float*>(tabv) + 4 * g), "f"(x), "f"(1.0f), "f"(0.0f), "f"(0.0f) : "memory");
      }
    }
  }
}

template <class L> static double time_ms(L launch, int warm = 1, int reps = 3) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < warm; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(a); launch(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  int lg = argc > 1 ? atoi(argv[1]) : 26;
  const long long G = argc > 2 ? atoll(argv[2]) : 100000;      // even
  long long N = 1ll << lg;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  long long* idx; float* a; double* tab; double* tab2;
  CK(cudaMalloc(&idx, N * 8)); CK(cudaMalloc(&a, N * 4));
  CK(cudaMalloc(&tab, G * 8)); CK(cudaMalloc(&tab2, G * 8));
  gen_uniform<<<sms * 8, 256>>>(idx, a, N, G, 100);
  CK(cudaDeviceSynchronize());
  std::vector<double> h1(G), h2(G);
  CK(cudaMemset(tab, 0, G * 8));
  CK(cudaFuncSetAttribute((const void*)k_tma_red<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 4 * 16));
  k_tma_red<0, 4><<<sms, 1024, 1024 * 4 * 16>>>(idx, a, N, tab);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h1.data(), tab, G * 8, cudaMemcpyDeviceToHost));
#define RUN(MODE, DEPTH)                                                                                              \
  {                                                                                                                   \
    const int smem = 1024 * DEPTH * 16;                                                                               \
    CK(cudaFuncSetAttribute((const void*)k_tma_red<MODE, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    CK(cudaMemset(tab2, 0, G * 8));                                                                                   \
    k_tma_red<MODE, DEPTH><<<sms, 1024, smem>>>(idx, a, N, tab2);                                                      \
    CK(cudaDeviceSynchronize());                                                                                      \
    CK(cudaMemcpy(h2.data(), tab2, G * 8, cudaMemcpyDeviceToHost));                                                   \
    const bool same = memcmp(h1.data(), h2.data(), G * 8) == 0;                                                       \
    double ms = time_ms([&] { k_tma_red<MODE, DEPTH><<<sms, 1024, smem>>>(idx, a, N, tab2); });                        \
    printf("{\"mode\": \"%s\", \"slots_per_thread\": %d, \"N\": %lld, \"G\": %lld, \"same_as_red\": %s, \"ms\": %.3f, \"g_updates_s\": %.1f}\n", \
           MODE == 0 ? "red.global.add.f64" : MODE == 1 ? "cp.reduce.async.bulk 16 B" : "half RED / half bulk (lane parity)",   \
           DEPTH, N, G, same ? "true" : "false", ms, N / ms * 1e-6);                                                  \
    fflush(stdout);                                                                                                   \
  }
  RUN(0, 4) RUN(1, 1) RUN(1, 2) RUN(1, 4) RUN(1, 8) RUN(2, 4) RUN(2, 8)
  {
    void* wide; CK(cudaMalloc(&wide, G * 16)); CK(cudaMemset(wide, 0, G * 16));
    const char* names[4] = {"two scalar REDs (f64 sum + u64 count)", "one red.global.add.f32", "one red.global.add.v2.f32", "one red.global.add.v4.f32"};
    for (int ctas : {111, 148}) {
      double ms[4];
      ms[0] = time_ms([&] { k_wide_red<0><<<ctas, 1024>>>(idx, a, N, wide); });
      ms[1] = time_ms([&] { k_wide_red<1><<<ctas, 1024>>>(idx, a, N, wide); });
      ms[2] = time_ms([&] { k_wide_red<2><<<ctas, 1024>>>(idx, a, N, wide); });
      ms[3] = time_ms([&] { k_wide_red<4><<<ctas, 1024>>>(idx, a, N, wide); });
      for (int i = 0; i < 4; ++i) {
        printf("{\"mode\": \"%s\", \"ctas\": %d, \"N\": %lld, \"G\": %lld, \"ms\": %.3f, \"g_elements_s\": %.1f}\n", names[i], ctas, N, G, ms[i], N / ms[i] * 1e-6);
      }
      fflush(stdout);
    }
  }
  // who bounds the scattered RED: the SM that sends it or the L2 / fabric that executes it?  The same N elements from fewer CTAs
  // (one CTA per SM): a sender-side bound scales with the number of SMs, a receiver-side bound does not
  for (int ctas : {18, 37, 55, 74, 92, 111, 129, 148}) {
    if (ctas > sms) continue;
    double ms = time_ms([&] { k_tma_red<0, 4><<<ctas, 1024, 1024 * 4 * 16>>>(idx, a, N, tab2); });
    printf("{\"mode\": \"red.global.add.f64, grid sweep\", \"ctas\": %d, \"N\": %lld, \"G\": %lld, \"ms\": %.3f, \"g_updates_s\": %.1f, \"red_per_clk_per_sm\": %.3f}\n",
           ctas, N, G, ms, N / ms * 1e-6, N / ms * 1e-6 / ctas / 1.965);
    fflush(stdout);
  }
  return 0;
}
