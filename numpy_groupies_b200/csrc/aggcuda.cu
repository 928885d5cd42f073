// aggcuda.cu -- extern "C" entry points of libaggcuda.so (see include/aggcuda.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC
#include <cstdio>
#include <cstring>
#include <type_traits>
#include "agc_common.cuh"
#include "agc_generic.cuh"
#include "agc_small.cuh"
#include "agc_fast.cuh"
#include <emmintrin.h>
#include "agc_dispatch.cuh"
#include "agc_scan.cuh"
#include "agc_labels.cuh"
#include "agc_dist.cuh"

namespace {

thread_local char g_err[512] = "";
int fail(int code, const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); return code; }
int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "CUDA error at %s: %s", where, cudaGetErrorString(e));
  return -100;
}
#define AGC_CUDA_OK(where) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return cuda_fail(e_, where); } while (0)

int sm_count() { return agc::device_sms(); }                  // of the CURRENT device

// measurement hook: a ring of event pairs per device (events belong to the device they were created on), one pair per
// agc_aggregate call -- a timed loop can run without reading anything back and collect the durations afterwards
bool g_profile = false;
constexpr int PROF_RING = 64;
struct ProfEvents { cudaEvent_t e0[PROF_RING] = {}, e1[PROF_RING] = {}; long long calls = 0; };
ProfEvents g_prof[agc::AGC_MAX_DEVICES];
int g_prof_last_dev = 0;
void prof_begin(cudaStream_t st) {
  if (!g_profile) return;
  ProfEvents& pe = g_prof[agc::current_device()];
  const int k = (int)(pe.calls % PROF_RING);
  if (!pe.e0[k]) { cudaEventCreate(&pe.e0[k]); cudaEventCreate(&pe.e1[k]); }
  cudaEventRecord(pe.e0[k], st);
}
void prof_end(cudaStream_t st) {
  if (!g_profile) return;
  const int dev = agc::current_device();
  ProfEvents& pe = g_prof[dev];
  cudaEventRecord(pe.e1[pe.calls % PROF_RING], st); ++pe.calls; g_prof_last_dev = dev;
}

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Layout { int64_t t0, t1, t2, b0, b1, extra, total; };
// table carve-up shared by agc_workspace_bytes / agc_partials / agc_aggregate
Layout reduce_layout(int64_t size) {
  Layout L; int64_t s8 = align_up(size * 8, 256), s1 = align_up(size, 256);
  L.t0 = 0; L.t1 = s8; L.t2 = 2 * s8; L.b0 = 3 * s8; L.b1 = 3 * s8 + s1; L.extra = 3 * s8 + 2 * s1;
  L.total = L.extra + align_up((int64_t)agc::PROBE_SCRATCH_WORDS * 4, 256);      // extra: the label probe's histograms
  return L;
}
agc::Tables tables_at(void* ws, const Layout& L) {
  char* b = (char*)ws; agc::Tables t;
  t.t0f = (double*)(b + L.t0); t.t0i = (long long*)(b + L.t0); t.t1 = (long long*)(b + L.t1);
  t.t2 = (double*)(b + L.t2); t.b0 = (unsigned char*)(b + L.b0); t.b1 = (unsigned char*)(b + L.b1);
  return t;
}

bool is_reduction(int op) { return (op >= AGC_OP_SUM && op <= AGC_OP_ANYNAN) || op == AGC_OP_MINMAX; }
bool is_two_pass(int op) { return op == AGC_OP_VAR || op == AGC_OP_ARGMAX || op == AGC_OP_ARGMIN; }

int check_request(const agc_request_t* r) {
  if (!r) return fail(-1, "null request");
  if (r->op < 0 || r->op >= AGC_OP_COUNT) return fail(-2, "unknown op");
  if (r->a_dtype < 0 || r->a_dtype >= AGC_DT_COUNT) return fail(-3, "bad a_dtype");
  if (r->out_dtype < 0 || r->out_dtype >= AGC_DT_COUNT) return fail(-3, "bad out_dtype");
  if (r->index.dtype < 0 || r->index.dtype > AGC_DT_U64) return fail(-3, "labels must have an integer dtype");
  if (r->n < 0 || r->size < 0) return fail(-4, "negative n or size");
  if (r->index.form < 0 || r->index.form > AGC_FORM_MULTI) return fail(-5, "bad index form");
  if (r->index.form != AGC_FORM_FLAT && (r->index.ndim < 1 || r->index.ndim > AGC_MAX_NDIM)) return fail(-5, "bad ndim");
  if (r->n > 0 && !r->index.labels) return fail(-6, "null labels");
  if (!r->status) return fail(-6, "null status");
  if (r->op == AGC_OP_MINMAX && !(r->flags & AGC_F_ACC_ONLY)) return fail(-2, "AGC_OP_MINMAX only accumulates (AGC_F_ACC_ONLY); finalise with op MAX / MIN");
  if (!r->out && !(r->flags & AGC_F_ACC_ONLY) && (r->size > 0 || !is_reduction(r->op))) return fail(-6, "null out");
  return 0;
}

int grid_for(int64_t work, int threads, int per_sm) {
  int64_t blocks = (work + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <class T>
void launch_accumulate(const agc::AccParams& p, cudaStream_t st) {
  if (p.n == 0) return;
  AGC_COUNT_LAUNCH(1), agc::k_accumulate<T><<<grid_for(p.n, 256, 16), 256, 0, st>>>(p);
}
void launch_accumulate_cls(const agc::AccParams& p, cudaStream_t st) {
  if (p.n == 0) return;
  if (!(p.flags & AGC_F_NO_FAST) && agc::small_path_ok(p.op, p.pass, p.n, p.size)) {       // few groups: per-CTA shared-memory copies
    switch (p.a_dtype) {
      case AGC_DT_F32: agc::launch_accumulate_small<float>(p, st, sm_count()); break;
      case AGC_DT_F64: agc::launch_accumulate_small<double>(p, st, sm_count()); break;
      case AGC_DT_U64: agc::launch_accumulate_small<unsigned long long>(p, st, sm_count()); break;
      default: agc::launch_accumulate_small<long long>(p, st, sm_count()); break;
    }
    return;
  }
  switch (p.a_dtype) {
    case AGC_DT_F32: launch_accumulate<float>(p, st); break;
    case AGC_DT_F64: launch_accumulate<double>(p, st); break;
    case AGC_DT_U64: launch_accumulate<unsigned long long>(p, st); break;
    default: launch_accumulate<long long>(p, st); break;
  }
}

}  // namespace

extern "C" {

int agc_abi_version(void) { return AGC_ABI_VERSION; }
const char* agc_last_error(void) { return g_err; }

int64_t agc_workspace_bytes(const agc_request_t* r) {
  if (check_request(r)) return -1;
  if (is_reduction(r->op)) return reduce_layout(r->size).total + 256;
  return agc::scan_workspace_bytes(r->op, r->n, r->size, r->a_dtype) + 256;
}

int agc_partials(const agc_request_t* r, agc_partial_t* out, int max_out) {
  if (check_request(r)) return -1;
  if (!is_reduction(r->op)) return fail(-7, "N->N operations have no mergeable partial tables");
  Layout L = reduce_layout(r->size);
  agc_partial_t tmp[5]; int k = 0;
  const bool fl = agc::acc_is_float(r->op, r->a_dtype, r->out_dtype);
  auto add = [&](int64_t off, int dt, int red) { tmp[k].offset_bytes = off; tmp[k].count = r->size; tmp[k].dtype = dt; tmp[k].reduce = red; ++k; };
  const bool pass2 = r->flags & AGC_F_PASS2;
  switch (r->op) {
    // the "touched" flags only travel when the finalise step reads them (fill_value differs from the identity)
    case AGC_OP_SUM: case AGC_OP_SUMSQ: add(L.t0, fl ? AGC_DT_F64 : AGC_DT_I64, 0); if (r->flags & AGC_F_NEED_TOUCHED) add(L.b1, AGC_DT_U8, 1); break;
    case AGC_OP_PROD: add(L.t0, fl ? AGC_DT_F64 : AGC_DT_I64, 3); if (r->flags & AGC_F_NEED_TOUCHED) add(L.b1, AGC_DT_U8, 1); break;
    case AGC_OP_LEN: add(L.t1, AGC_DT_I64, 0); break;
    case AGC_OP_MEAN: add(L.t0, AGC_DT_F64, 0); add(L.t1, AGC_DT_I64, 0); break;
    case AGC_OP_VAR: if (pass2) add(L.t2, AGC_DT_F64, 0); else { add(L.t0, AGC_DT_F64, 0); add(L.t1, AGC_DT_I64, 0); } break;
    // float max / min tell "empty" by the key itself; integer values need the flags
    case AGC_OP_MAX: add(L.t0, AGC_DT_I64, 1); if (!agc::dtype_is_float(r->a_dtype)) add(L.b1, AGC_DT_U8, 1); break;
    case AGC_OP_MIN: add(L.t0, AGC_DT_I64, 2); if (!agc::dtype_is_float(r->a_dtype)) add(L.b1, AGC_DT_U8, 1); break;
    case AGC_OP_MINMAX: add(L.t0, AGC_DT_I64, 1); add(L.t2, AGC_DT_I64, 2); if (!agc::dtype_is_float(r->a_dtype)) add(L.b1, AGC_DT_U8, 1); break;
    case AGC_OP_ARGMAX: if (pass2) add(L.t1, AGC_DT_I64, 2); else add(L.t0, AGC_DT_I64, 1); break;
    case AGC_OP_ARGMIN: if (pass2) add(L.t1, AGC_DT_I64, 2); else add(L.t0, AGC_DT_I64, 2); break;
    case AGC_OP_ANY: case AGC_OP_ANYNAN: add(L.b0, AGC_DT_U8, 1); add(L.b1, AGC_DT_U8, 1); break;
    case AGC_OP_ALL: case AGC_OP_ALLNAN: add(L.b0, AGC_DT_U8, 2); add(L.b1, AGC_DT_U8, 1); break;
    case AGC_OP_FIRST: add(L.t1, AGC_DT_I64, 2); break;
    case AGC_OP_LAST: add(L.t1, AGC_DT_I64, 1); break;
    default: break;
  }
  for (int i = 0; i < k && i < max_out; ++i) out[i] = tmp[i];
  return k;
}

int agc_aggregate(const agc_request_t* r, void* stream) {
  int rc = check_request(r);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!is_reduction(r->op)) {
    rc = agc::run_scan_op(r, st, sm_count());
    if (rc) return fail(rc, agc::scan_error());
    AGC_CUDA_OK("N->N op");
    return 0;
  }
  if (r->workspace_bytes < reduce_layout(r->size).total) return fail(-8, "workspace too small");
  if (!r->workspace && r->size > 0) return fail(-6, "null workspace");
  const Layout L = reduce_layout(r->size);
  agc::Tables t = tables_at(r->workspace, L);
  const bool acc = !(r->flags & AGC_F_FIN_ONLY), fin = !(r->flags & AGC_F_ACC_ONLY);
  const bool pass2_only = (r->flags & AGC_F_PASS2) != 0;

  agc::AccParams p;
  memset(&p, 0, sizeof(p));
  p.op = r->op; p.flags = r->flags; p.a_dtype = r->a_dtype; p.out_dtype = r->out_dtype; p.pass = 1;
  p.a = r->a; p.a_sf = r->a_scalar_f64; p.a_si = r->a_scalar_i64;
  p.ix.ix = r->index; p.ix.size = r->size; p.ix.n = r->n;
  p.n = r->n; p.size = r->size; p.index_base = r->index_base; p.t = t; p.status = r->status;

  if (acc) {
    bool done = false;
    if (!pass2_only) {
      if (!(r->flags & AGC_F_NO_FAST) && !(r->flags & AGC_F_NO_INIT) && r->size > 0 && r->n > 0) {
        // specialised shared-memory kernels (they initialise + write the same global tables)
        prof_begin(st);
        int frc = agc::try_fast_path(r, t, st, sm_count(), (unsigned*)((char*)r->workspace + L.extra));
        if (frc < 0) return fail(frc, "fast path launch failed");
        done = frc > 0;
        if (done) prof_end(st);
        AGC_CUDA_OK("fast path");
      }
      if (!done) {
        if (!(r->flags & AGC_F_NO_INIT) && r->size > 0) {
          AGC_COUNT_LAUNCH(1), agc::k_init_tables<<<grid_for(r->size, 256, 8), 256, 0, st>>>(t, r->size, r->op, r->a_dtype, r->out_dtype, 0);
          AGC_CUDA_OK("init");
        }
        prof_begin(st);
        launch_accumulate_cls(p, st);
        prof_end(st);
        AGC_CUDA_OK("accumulate");
      }
    }
    if (is_two_pass(r->op) && (pass2_only || fin)) {
      if (r->size > 0) {
        if (r->op == AGC_OP_VAR) {
          AGC_COUNT_LAUNCH(1), agc::k_means_inplace<<<grid_for(r->size, 256, 8), 256, 0, st>>>(t, r->size);
          if (pass2_only) AGC_COUNT_LAUNCH(1), agc::k_init_tables<<<grid_for(r->size, 256, 8), 256, 0, st>>>(t, r->size, r->op, r->a_dtype, r->out_dtype, 1);
        } else if (pass2_only) {
          AGC_COUNT_LAUNCH(1), agc::k_init_tables<<<grid_for(r->size, 256, 8), 256, 0, st>>>(t, r->size, r->op, r->a_dtype, r->out_dtype, 1);
        }
      }
      p.pass = 2;
      int frc2 = 0;
      if (r->op == AGC_OP_VAR && !(r->flags & AGC_F_NO_FAST) && r->n > 0 && r->size > 0) {
        frc2 = agc::try_fast_pass2(r, t, st, sm_count());
        if (frc2 < 0) return fail(frc2, "fast second pass launch failed");
      }
      if (frc2 == 0) launch_accumulate_cls(p, st);
      AGC_CUDA_OK("second pass");
    }
  }
  if (fin && r->size > 0) {
    agc::FinParams f;
    memset(&f, 0, sizeof(f));
    f.op = r->op; f.flags = r->flags; f.a_dtype = r->a_dtype; f.out_dtype = r->out_dtype; f.a = r->a; f.out = r->out;
    f.fill_f = r->fill_f64; f.fill_i = r->fill_i64; f.ddof = r->ddof; f.size = r->size; f.index_base = r->index_base;
    f.arg_stride = r->arg_stride; f.arg_len = r->arg_len; f.n = r->n; f.t = t;
    AGC_COUNT_LAUNCH(1), agc::k_finalize<<<grid_for(r->size, 256, 8), 256, 0, st>>>(f);
    AGC_CUDA_OK("finalize");
  }
  return 0;
}

int agc_profile_enable(int on) { g_profile = on != 0; for (auto& pe : g_prof) pe.calls = 0; return 0; }
int agc_host_narrow_labels(const void* src, int32_t src_dtype, int64_t n, int32_t* dst) {
  if (src_dtype != AGC_DT_I64 && src_dtype != AGC_DT_U64) return fail(-1, "agc_host_narrow_labels: 8-byte integer labels only");
  const unsigned long long* s = (const unsigned long long*)src;
  int64_t i = 0;
  auto one = [&](int64_t k) { const unsigned long long v = s[k]; dst[k] = (v >> 31) ? -1 : (int32_t)v; };
  // The staging threads are bound by the host's memory system, not by arithmetic: non-temporal stores keep the destination
  // (a page-locked buffer the DMA engine reads next) out of the caches and spare the read-for-ownership of every line --
  // 12 instead of 16 bytes of DRAM traffic per label.
  while (i < n && ((uintptr_t)(dst + i) & 15)) one(i++);
  const __m128i zero = _mm_setzero_si128();
  for (; i + 8 <= n; i += 8) {
    const __m128 a0 = _mm_castsi128_ps(_mm_loadu_si128((const __m128i*)(s + i))), a1 = _mm_castsi128_ps(_mm_loadu_si128((const __m128i*)(s + i + 2)));
    const __m128 b0 = _mm_castsi128_ps(_mm_loadu_si128((const __m128i*)(s + i + 4))), b1 = _mm_castsi128_ps(_mm_loadu_si128((const __m128i*)(s + i + 6)));
    const __m128i lo0 = _mm_castps_si128(_mm_shuffle_ps(a0, a1, 0x88)), hi0 = _mm_castps_si128(_mm_shuffle_ps(a0, a1, 0xDD));
    const __m128i lo1 = _mm_castps_si128(_mm_shuffle_ps(b0, b1, 0x88)), hi1 = _mm_castps_si128(_mm_shuffle_ps(b0, b1, 0xDD));
    const __m128i ok0 = _mm_cmpeq_epi32(_mm_or_si128(hi0, _mm_srai_epi32(lo0, 31)), zero);      // label in [0, 2^31)
    const __m128i ok1 = _mm_cmpeq_epi32(_mm_or_si128(hi1, _mm_srai_epi32(lo1, 31)), zero);
    _mm_stream_si128((__m128i*)(dst + i), _mm_or_si128(_mm_and_si128(lo0, ok0), _mm_andnot_si128(ok0, _mm_set1_epi32(-1))));
    _mm_stream_si128((__m128i*)(dst + i + 4), _mm_or_si128(_mm_and_si128(lo1, ok1), _mm_andnot_si128(ok1, _mm_set1_epi32(-1))));
  }
  for (; i < n; ++i) one(i);
  _mm_sfence();
  return 0;
}
int agc_host_copy_stream(const void* src, void* dst, int64_t bytes) {
  const char* s = (const char*)src; char* d = (char*)dst;
  int64_t i = 0;
  while (i < bytes && ((uintptr_t)(d + i) & 15)) { d[i] = s[i]; ++i; }
  for (; i + 64 <= bytes; i += 64) {
    const __m128i v0 = _mm_loadu_si128((const __m128i*)(s + i)), v1 = _mm_loadu_si128((const __m128i*)(s + i + 16));
    const __m128i v2 = _mm_loadu_si128((const __m128i*)(s + i + 32)), v3 = _mm_loadu_si128((const __m128i*)(s + i + 48));
    _mm_stream_si128((__m128i*)(d + i), v0); _mm_stream_si128((__m128i*)(d + i + 16), v1);
    _mm_stream_si128((__m128i*)(d + i + 32), v2); _mm_stream_si128((__m128i*)(d + i + 48), v3);
  }
  for (; i < bytes; ++i) d[i] = s[i];
  _mm_sfence();
  return 0;
}
float agc_profile_ms(int back) {
  ProfEvents& pe = g_prof[g_prof_last_dev];
  if (back < 0 || back >= PROF_RING || pe.calls - 1 - back < 0) return -1.f;
  const int k = (int)((pe.calls - 1 - back) % PROF_RING);
  float ms = -1.f;
  if (cudaEventSynchronize(pe.e1[k]) != cudaSuccess) return -1.f;
  cudaEventElapsedTime(&ms, pe.e0[k], pe.e1[k]);
  return ms;
}
float agc_profile_last_ms(void) { return agc_profile_ms(0); }
int64_t agc_launch_count(void) { return agc::g_launches.load(); }
int agc_tuning(int key, int value) {
  agc::tuning_defaults();
  if (key == agc::TUNE_SCAN_CHAIN) { if (value >= 0) agc::g_scan_chain = value ? 1 : 0; return agc::scan_chain_enabled() ? 1 : 0; }
  if (key < 0 || key >= agc::TUNE_COUNT) return -1;
  if (value >= 0) agc::g_tune[key] = value;
  return agc::g_tune[key];
}

int agc_label_minmax(const void* labels, int32_t dtype, int64_t n, int64_t* out_minmax, void* stream) {
  if (n <= 0 || !labels || !out_minmax) return fail(-6, "agc_label_minmax: empty or null input");
  if (dtype < 0 || dtype > AGC_DT_U64) return fail(-3, "labels must have an integer dtype");
  cudaStream_t st = (cudaStream_t)stream;
  AGC_COUNT_LAUNCH(1), agc::k_minmax_init<<<1, 1, 0, st>>>((long long*)out_minmax);
  AGC_COUNT_LAUNCH(1), agc::k_label_minmax<<<grid_for(n, 256, 8), 256, 0, st>>>(labels, dtype, n, (long long*)out_minmax);
  AGC_CUDA_OK("label_minmax");
  return 0;
}

int agc_unpack(const void* table, int32_t elem_bytes, int64_t size, const void* labels, int32_t dtype, int64_t n,
               void* out, int32_t* status, void* stream) {
  if (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8) return fail(-3, "elem_bytes must be 1/2/4/8");
  if (n == 0) return 0;
  if (!table || !labels || !out || !status) return fail(-6, "agc_unpack: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  AGC_COUNT_LAUNCH(1), agc::k_unpack<<<grid_for(n, 256, 16), 256, 0, st>>>(table, elem_bytes, size, labels, dtype, n, out, status);
  AGC_CUDA_OK("unpack");
  return 0;
}

int agc_convert_f16(const void* src, void* dst, int64_t n, int32_t to_half, void* stream) {
  if (n < 0) return fail(-3, "agc_convert_f16: negative length");
  if (n == 0) return 0;
  if (!src || !dst) return fail(-6, "agc_convert_f16: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (to_half) AGC_COUNT_LAUNCH(1), agc::k_convert_f16<true><<<grid_for((n + 7) >> 3, 256, 8), 256, 0, st>>>(src, dst, n);
  else AGC_COUNT_LAUNCH(1), agc::k_convert_f16<false><<<grid_for((n + 7) >> 3, 256, 8), 256, 0, st>>>(src, dst, n);
  AGC_CUDA_OK("convert_f16");
  return 0;
}

int agc_pack_tables(const agc_pack_seg_t* segs, int32_t nseg, void* packed, int32_t carrier_dtype, int32_t unpack, void* stream) {
  if (nseg < 1 || nseg > agc::PK_MAX_SEGS || !segs || !packed) return fail(-6, "agc_pack_tables: bad segment list");
  if (carrier_dtype != AGC_DT_F64 && carrier_dtype != AGC_DT_I64) return fail(-3, "agc_pack_tables: the carrier is float64 or int64");
  agc::PackParams p;
  memset(&p, 0, sizeof(p));
  long long total = 0;
  for (int i = 0; i < nseg; ++i) {
    if (segs[i].count < 0 || (segs[i].count > 0 && !segs[i].table)) return fail(-6, "agc_pack_tables: null table");
    const int dt = segs[i].dtype;
    if (dt != AGC_DT_F64 && dt != AGC_DT_I64 && dt != AGC_DT_I32 && dt != AGC_DT_U8 && dt != AGC_DT_BOOL) return fail(-3, "agc_pack_tables: table dtype");
    p.seg[i].table = segs[i].table; p.seg[i].count = segs[i].count; p.seg[i].dtype = dt; p.seg[i].xform = segs[i].xform;
    p.start[i] = total; total += segs[i].count;
  }
  p.start[nseg] = total; p.nseg = nseg; p.carrier = carrier_dtype; p.packed = packed; p.unpack = unpack != 0;
  if (total == 0) return 0;
  AGC_COUNT_LAUNCH(1), agc::k_pack_tables<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(p);
  AGC_CUDA_OK("pack_tables");
  return 0;
}

int agc_rank_fold(const void* gathered, int32_t rank, int64_t size, int32_t dtype, int32_t reduce, void* carry, void* stream) {
  if (dtype != AGC_DT_F64 && dtype != AGC_DT_I64) return fail(-3, "agc_rank_fold: tables are float64 or int64");
  if (reduce < 0 || reduce > 3 || rank < 0) return fail(-2, "agc_rank_fold: bad reduce / rank");
  if (size <= 0) return 0;
  if (!gathered || !carry) return fail(-6, "agc_rank_fold: null pointer");
  AGC_COUNT_LAUNCH(1), agc::k_rank_fold<<<grid_for(size, 256, 8), 256, 0, (cudaStream_t)stream>>>(gathered, rank, size, dtype == AGC_DT_F64, reduce, carry);
  AGC_CUDA_OK("rank_fold");
  return 0;
}

int agc_apply_carry(int32_t op, const void* labels, int32_t label_dtype, int64_t n, int64_t size, void* out, int32_t out_dtype,
                    int32_t a_dtype, const void* carry, int32_t carry_dtype, int32_t* status, void* stream) {
  if (op < AGC_OP_CUMSUM || op > AGC_OP_CUMMIN) return fail(-2, "agc_apply_carry: op must be a running function");
  if (n <= 0 || size <= 0) return 0;
  if (!labels || !out || !carry) return fail(-6, "agc_apply_carry: null pointer");
  agc::CarryParams p;
  p.labels = labels; p.ldtype = label_dtype; p.n = n; p.size = size; p.out = out; p.out_dtype = out_dtype;
  p.carry = carry; p.acc_f64 = carry_dtype == AGC_DT_F64;
  p.op = op == AGC_OP_CUMSUM ? 0 : (op == AGC_OP_CUMMAX ? 1 : (op == AGC_OP_CUMMIN ? 2 : 3));
  p.a_float = agc::dtype_is_float(a_dtype); p.a_u64 = a_dtype == AGC_DT_U64; p.status = status;
  AGC_COUNT_LAUNCH(1), agc::k_apply_carry<<<grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(p);
  AGC_CUDA_OK("apply_carry");
  return 0;
}

int agc_ordered_accumulate(const agc_request_t* r, const int64_t* order, const int64_t* offsets, void* stream) {
  if (check_request(r)) return -1;
  if (!(r->op == AGC_OP_SUM || r->op == AGC_OP_SUMSQ || r->op == AGC_OP_MEAN || r->op == AGC_OP_VAR || r->op == AGC_OP_PROD))
    return fail(-20, "deterministic mode covers sum / sumofsquares / mean / var / std / prod");
  if (r->a_dtype != AGC_DT_F32 && r->a_dtype != AGC_DT_F64) return fail(-20, "deterministic mode: integer inputs are order-independent already; values must be float32 / float64");
  if (!r->a || !order || !offsets) return fail(-6, "agc_ordered_accumulate: null pointer");
  if (r->workspace_bytes < reduce_layout(r->size).total) return fail(-8, "workspace too small");
  if (r->size <= 0) return 0;
  agc::OrderedParams p;
  p.a = r->a; p.a_dtype = r->a_dtype; p.order = (const long long*)order; p.offsets = (const long long*)offsets; p.size = r->size;
  p.t = tables_at(r->workspace, reduce_layout(r->size)); p.op = r->op; p.flags = r->flags;
  AGC_COUNT_LAUNCH(1), agc::k_ordered_accumulate<<<grid_for(r->size * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(p);
  AGC_CUDA_OK("ordered_accumulate");
  return 0;
}

int agc_owner_pack(const agc_request_t* r, int64_t* carrier, int32_t unpack, void* stream) {
  if (check_request(r)) return -1;
  if (r->op != AGC_OP_FIRST && r->op != AGC_OP_LAST) return fail(-2, "agc_owner_pack: first / last only");
  if (r->size <= 0) return 0;
  if (!carrier || !r->out || !r->workspace) return fail(-6, "agc_owner_pack: null pointer");
  const Layout L = reduce_layout(r->size);
  agc::Tables t = tables_at(r->workspace, L);
  AGC_COUNT_LAUNCH(1), agc::k_owner_pack<<<grid_for(r->size, 256, 8), 256, 0, (cudaStream_t)stream>>>(r->out, agc::dtype_bytes(r->out_dtype), r->size, t.t1,
                                                                                                r->index_base, r->n, (long long*)carrier, unpack != 0);
  AGC_CUDA_OK("owner_pack");
  return 0;
}

int64_t agc_label_scan_workspace_bytes(int64_t n) { return agc::label_scan_workspace_bytes(n < 0 ? 0 : n); }

int agc_label_scan(int32_t mode, const void* x, int32_t x_dtype, int64_t n, void* out, int32_t out_dtype, int64_t* total,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  if (mode < 0 || mode > agc::LB_CUMSUM) return fail(-2, "agc_label_scan: unknown mode");
  if (x_dtype < 0 || x_dtype > AGC_DT_U64) return fail(-3, "agc_label_scan: x must have a bool / integer dtype");
  if (n < 0) return fail(-4, "negative n");
  if (n > 0 && (!x || !workspace || (mode != agc::LB_STEP_COUNT && !out))) return fail(-6, "agc_label_scan: null pointer");
  if (workspace_bytes < agc::label_scan_workspace_bytes(n)) return fail(-8, "workspace too small");
  if (mode == agc::LB_KEEP && (out_dtype < 0 || out_dtype > AGC_DT_U64)) return fail(-3, "relabel table needs an integer dtype");
  agc::LabelScanParams p;
  p.x = x; p.x_dtype = x_dtype; p.n = n; p.out = out; p.out_dtype = out_dtype; p.total = (long long*)total; p.desc = nullptr;
  int rc = agc::run_label_scan(mode, p, workspace, (cudaStream_t)stream);
  if (rc) return fail(rc, "agc_label_scan failed");
  AGC_CUDA_OK("label_scan");
  return 0;
}

int agc_mark_present(const void* labels, int32_t dtype, int64_t n, int64_t size, void* keep, int32_t* status, void* stream) {
  if (dtype < 0 || dtype > AGC_DT_U64) return fail(-3, "labels must have an integer dtype");
  if (n <= 0) return 0;
  if (!labels || !keep || !status) return fail(-6, "agc_mark_present: null pointer");
  AGC_COUNT_LAUNCH(1), agc::k_mark_present<<<grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(labels, dtype, n, size, (unsigned char*)keep, status);
  AGC_CUDA_OK("mark_present");
  return 0;
}

int agc_multi_arange(const int64_t* incl, int64_t m, int64_t total, int64_t* out, void* stream) {
  if (total <= 0 || m <= 0) return 0;
  if (!incl || !out) return fail(-6, "agc_multi_arange: null pointer");
  AGC_COUNT_LAUNCH(1), agc::k_multi_arange<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>((const long long*)incl, m, total, (long long*)out);
  AGC_CUDA_OK("multi_arange");
  return 0;
}

int64_t agc_group_order_workspace_bytes(int64_t n, int64_t size) { return agc::sort_workspace_bytes(n, size) + 256; }

int agc_group_order(const agc_index_t* index, int64_t n, int64_t size, int64_t* order, int64_t* offsets,
                    void* workspace, int64_t workspace_bytes, int32_t* status, void* stream) {
  if (!index || !order || !status) return fail(-6, "agc_group_order: null pointer");
  if (workspace_bytes < agc::sort_workspace_bytes(n, size)) return fail(-8, "workspace too small");
  int rc = agc::group_order(*index, n, size, (long long*)order, (long long*)offsets, workspace, status, (cudaStream_t)stream, sm_count());
  if (rc) return fail(rc, agc::scan_error());
  AGC_CUDA_OK("group_order");
  return 0;
}

}  // extern "C"
