// team_probe.cu -- stand-alone correctness + timing harness for k_team (agc_team.cuh), used while the kernel was tuned.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o ubench/team_probe ubench/team_probe.cu
//   ubench/team_probe C xport mode log2n groups [smem_budget] [cap] [reps] [max_clusters] [dist: 0 uniform, 1 skewed]
// Prints one JSON line: config, max relative error of the sums against plain f64 global atomics, count equality, ms.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "../numpy_groupies_b200/csrc/agc_team.cuh"

using namespace agc;

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
__global__ void k_gen(long long* labels, float* a, long long n, unsigned long long G, unsigned long long seed, int dist) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long h = mix64(seed + 2 * (unsigned long long)i), h2 = mix64(seed + 2 * (unsigned long long)i + 1);
    unsigned long long g = h % G;
    if (dist == 1 && (h2 & 7) == 0) g = (h >> 20) % 16;            // 1/8 of the elements on 16 hot groups
    labels[i] = (long long)g;
    float x = (float)(h2 >> 40) * 5.9604644775390625e-08f;          // multiples of 2^-24 in [0, 1), like torch.rand
    if (dist == 2) x = (x - 0.5f) * ((h2 & 0xff00) == 0 ? 1.0e-6f : 1.0f);   // signed, a few tiny values
    a[i] = x;
  }
}
__global__ void k_ref(const long long* labels, const float* a, long long n, double* sums, unsigned long long* cnt) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    atomicAdd(sums + labels[i], (double)a[i]);
    atomicAdd(cnt + labels[i], 1ull);
  }
}
__global__ void k_cmp(const double* s0, const double* s1, const unsigned long long* c0, const long long* c1, long long G, int check_cnt, double* out) {
  // out[0] = max |s0 - s1| / max(|s0|, tiny); out[1] = number of count mismatches
  const long long stride = (long long)gridDim.x * blockDim.x;
  double worst = 0.0; unsigned long long mism = 0;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < G; g += stride) {
    const double d = fabs(s0[g] - s1[g]) / fmax(fabs(s0[g]), 1e-30);
    worst = fmax(worst, d);
    if (check_cnt && c0[g] != (unsigned long long)c1[g]) ++mism;
  }
  atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(worst));   // non-negative doubles order like integers
  if (mism) atomicAdd((unsigned long long*)(out + 1), mism);
}


// ---- transport-only microbenchmark: the k_team exchange protocol with synthetic records, no global loads, no table ----
// every warp sends K records to each peer per phase (K/32 contiguous stores per lane) and sums what it receives
template <int C, int XPORT>
__global__ void __launch_bounds__(TEAM_THREADS, 1) k_xport(int phases, int K, int cap, unsigned long long* sink) {
  extern __shared__ __align__(16) unsigned sw[];
  constexpr int P = C - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned me = dsm::ctarank();
  unsigned long long* inbox = (unsigned long long*)sw;
  unsigned long long* bars = inbox + (size_t)P * 2 * TEAM_WARPS * cap;
  for (int i = threadIdx.x; i < 2 * P * 2 * TEAM_WARPS; i += TEAM_THREADS) dsm::mbar_init(dsm::smem_u32(bars + i), 1);
  dsm::fence_mbar_init();
  dsm::cluster_sync();
  unsigned rbase[P];
  const unsigned sw_s = dsm::smem_u32(sw);
#pragma unroll
  for (int j = 0; j < P; ++j) rbase[j] = dsm::mapa(sw_s, (me + j + 1) & (C - 1)) - sw_s;
  constexpr unsigned BAR_B = TEAM_WARPS * 8;
  const unsigned buf_stride = (unsigned)(TEAM_WARPS * cap * 8);
  const unsigned in0 = dsm::smem_u32(inbox) + (unsigned)warp * cap * 8u;
  const unsigned full0 = dsm::smem_u32(bars) + (unsigned)warp * 8u;
  const unsigned empty0 = full0 + P * 2 * BAR_B;
  unsigned long long acc = 0;
  auto consume = [&](int ph) {
    const unsigned b = (unsigned)ph & 1u, par = (unsigned)(ph >> 1) & 1u;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      dsm::mbar_wait(full0 + (j * 2 + b) * BAR_B, par);
      const unsigned long long* reg = inbox + ((size_t)(j * 2 + b) * TEAM_WARPS + warp) * cap;
      const unsigned cnt = (unsigned)reg[0];
      for (unsigned i = lane; i < cnt; i += 32) acc += reg[1 + i];
      __syncwarp();
      if (lane == 0) dsm::remote_arrive_relaxed(rbase[P - 1 - j] + empty0 + (j * 2 + b) * BAR_B);
    }
  };
  for (int ph = 0; ph < phases; ++ph) {
    const unsigned b = (unsigned)ph & 1u;
    if (ph >= 2) {
      const unsigned par = (unsigned)((ph >> 1) - 1) & 1u;
#pragma unroll
      for (int j = 0; j < P; ++j) dsm::mbar_wait(empty0 + (j * 2 + b) * BAR_B, par);
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const unsigned dst = rbase[j] + in0 + (j * 2 + b) * buf_stride, rb = rbase[j] + full0 + (j * 2 + b) * BAR_B;
      for (int s = lane; s < K; s += 32) {
        if (XPORT == 0) dsm::st_async_rec(dst + (1u + s) * 8u, (unsigned)(ph + s), me, rb);
        else dsm::st_remote_rec(dst + (1u + s) * 8u, (unsigned)(ph + s), me);
      }
      if (XPORT != 0) __syncwarp();
      if (lane == 0) {
        if (XPORT == 0) { dsm::st_async_rec(dst, (unsigned)K, 0u, rb); dsm::remote_arrive_expect_tx(rb, 8u * (K + 1u)); }
        else { dsm::st_remote_rec(dst, (unsigned)K, 0u); dsm::remote_arrive_release(rb); }
      }
    }
    if (ph >= 1) consume(ph - 1);
  }
  if (phases >= 1) consume(phases - 1);
  dsm::cluster_sync();
  if (acc == 0x1234567ull) sink[0] = acc;
  // every warp received P * K records per phase, each {lo = ph + s, hi = sender}: a checksum proves delivery
  acc = __reduce_add_sync(0xffffffffu, (unsigned)(acc & 0xffffffffu)) ;
  if (lane == 0) atomicAdd(sink + 1, acc);
}
template <int C, int XPORT>
int run_xport(int phases, int K) {
  const int cap = ((K + 1) + 3) & ~3;
  const size_t smem = (size_t)(C - 1) * 2 * TEAM_WARPS * cap * 8 + (size_t)2 * (C - 1) * 2 * TEAM_WARPS * 8;
  auto kern = k_xport<C, XPORT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64);
  cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(TEAM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = 0; cfg.attrs = at; cfg.numAttrs = 1; cfg.gridDim = dim3(C);
  int ncl = 0;
  if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) != cudaSuccess || ncl < 1) { printf("{\"error\": \"occupancy\"}\n"); return 1; }
  cfg.gridDim = dim3(ncl * C);
  unsigned long long* sink; cudaMalloc(&sink, 16); cudaMemset(sink, 0, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaMemset(sink, 0, 16);
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, kern, phases, K, cap, sink);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  unsigned long long h[2]; cudaMemcpy(h, sink, 16, cudaMemcpyDeviceToHost);
  // expected checksum: per warp and phase, P peers x sum_{s<K} (ph + s), low 32 bits summed over lanes (mod 2^64 of 32-bit pieces is not exact: report only)
  const double recs = (double)ncl * C * TEAM_WARPS * (double)phases * (C - 1) * K;
  printf("{\"xport_bench\": 1, \"C\": %d, \"xport\": %d, \"K\": %d, \"phases\": %d, \"clusters\": %d, \"ms\": %.4f, \"grec_s\": %.1f, \"rec_per_clk_sm\": %.3f, \"checksum\": %llu}\n",
         C, XPORT, K, phases, ncl, best, recs / best * 1e-6, recs / (best * 1e-3) / 1.965e9 / (ncl * C), h[1]);
  return 0;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

template <int C, int MODE, int NS>
int launch(const TeamParams& p, size_t smem, int max_clusters) { return launch_team<C, long long, MODE, NS != 0>(p, smem, 0, max_clusters); }

int dispatch(int C, int mode, int xport, const TeamParams& p, size_t smem, int mc) {
#define ONE(C_, M_, X_) if (C == C_ && mode == M_ && xport == X_) return launch<C_, M_, X_>(p, smem, mc);
  ONE(2, 0, 0) ONE(2, 0, 1) ONE(2, 1, 0) ONE(2, 1, 1)
  ONE(4, 0, 0) ONE(4, 0, 1) ONE(4, 1, 0) ONE(4, 1, 1)
#undef ONE
  return -99;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "xport")) {          // team_probe xport C xport K phases
    const int C = atoi(argv[2]), X = atoi(argv[3]), K = atoi(argv[4]), ph = atoi(argv[5]);
    if (C == 2 && X == 0) return run_xport<2, 0>(ph, K);
    if (C == 2 && X == 1) return run_xport<2, 1>(ph, K);
    if (C == 4 && X == 0) return run_xport<4, 0>(ph, K);
    if (C == 4 && X == 1) return run_xport<4, 1>(ph, K);
    if (C == 8 && X == 0) return run_xport<8, 0>(ph, K);
    return 2;
  }
  if (argc < 6) { printf("usage: team_probe C xport mode log2n groups [smem_budget] [cap] [reps] [max_clusters] [dist]\n"); return 2; }
  const int C = atoi(argv[1]), xport = atoi(argv[2]), mode = atoi(argv[3]), log2n = atoi(argv[4]);
  const long long G = atoll(argv[5]);
  const long long budget = argc > 6 ? atoll(argv[6]) : 199552;
  int cap = argc > 7 ? atoi(argv[7]) : 0;
  const int reps = argc > 8 ? atoi(argv[8]) : 5;
  const int mc = argc > 9 ? atoi(argv[9]) : 0;
  const int dist = argc > 10 ? atoi(argv[10]) : 0;
  const int dbg = argc > 11 ? atoi(argv[11]) : 0;
  if (cap <= 0) cap = team_default_cap(C);
  const long long n = (1ll << log2n) + (log2n < 30 ? 3 : 0);                 // ragged tail on the small cases
  const size_t fixed = team_smem_bytes(C, mode, 0, cap);
  long long cov_need = G / C;                                                  // cov * C <= size (the kernel relies on it)
  long long cov_max = mode == TM_SUM ? (budget - (long long)fixed - 16) / 4 : (budget - (long long)fixed - 16) / 6;
  const int cov = (int)(cov_need < cov_max ? cov_need : cov_max);
  const size_t smem = team_smem_bytes(C, mode, cov, cap);

  long long* labels; float* a; double *s_ref, *s_out, *d_cmp; unsigned long long* c_ref; long long* c_out; unsigned char* b1; int* status;
  CK(cudaMalloc(&labels, n * 8)); CK(cudaMalloc(&a, n * 4));
  CK(cudaMalloc(&s_ref, G * 8)); CK(cudaMalloc(&s_out, G * 8)); CK(cudaMalloc(&c_ref, G * 8)); CK(cudaMalloc(&c_out, G * 8));
  CK(cudaMalloc(&b1, G)); CK(cudaMalloc(&status, 64)); CK(cudaMalloc(&d_cmp, 16));
  k_gen<<<148 * 8, 256>>>(labels, a, n, (unsigned long long)G, 1234567ull, dist);
  CK(cudaMemset(s_ref, 0, G * 8)); CK(cudaMemset(c_ref, 0, G * 8)); CK(cudaMemset(s_out, 0, G * 8)); CK(cudaMemset(c_out, 0, G * 8));
  CK(cudaMemset(b1, 0, G)); CK(cudaMemset(status, 0, 64)); CK(cudaMemset(d_cmp, 0, 16));
  k_ref<<<148 * 8, 256>>>(labels, a, n, s_ref, c_ref);
  CK(cudaDeviceSynchronize());

  TeamParams p;
  memset(&p, 0, sizeof(p));
  p.labels = labels; p.a = a; p.n = n; p.size = G;
  p.t.t0f = s_out; p.t.t0i = (long long*)s_out; p.t.t1 = c_out; p.t.b1 = b1; p.status = status;
  p.flags = 0; p.want_cnt = 1; p.want_b1 = 0; p.cov = cov; p.cap = cap; p.choose = nullptr; p.want = 0; (void)dbg;

  int ncl = dispatch(C, mode, xport, p, smem, mc);
  if (ncl < 0) { printf("{\"error\": \"launch rc %d\", \"smem\": %zu}\n", ncl, smem); return 1; }
  CK(cudaDeviceSynchronize());
  k_cmp<<<148, 256>>>(s_ref, s_out, c_ref, c_out, G, mode == TM_SUMCNT, d_cmp);
  double h_cmp[2]; int h_status[8];
  CK(cudaMemcpy(h_cmp, d_cmp, 16, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_status, status, 32, cudaMemcpyDeviceToHost));
  unsigned long long mism; memcpy(&mism, &h_cmp[1], 8);

  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f, tot = 0.f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    dispatch(C, mode, xport, p, smem, mc);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best; tot += ms;
  }
  printf("{\"C\": %d, \"xport\": %d, \"mode\": %d, \"log2n\": %d, \"groups\": %lld, \"dist\": %d, \"cov\": %d, \"cov_need\": %lld, \"cap\": %d, \"smem\": %zu, "
         "\"dbg\": %d, \"clusters\": %d, \"max_rel_err\": %.3e, \"count_mismatch\": %llu, \"bad\": %d, \"regime\": %d, \"ms_best\": %.4f, \"ms_mean\": %.4f, "
         "\"gelem_s\": %.1f, \"frac_6557\": %.4f}\n",
         C, xport, mode, log2n, G, dist, cov, cov_need, cap, smem, dbg, ncl, h_cmp[0], mism, h_status[0], h_status[3], best, tot / reps,
         (double)n / best * 1e-6, ((double)n * 12 + G * 4) / best * 1e-6 / 6557.1);
  return 0;
}
