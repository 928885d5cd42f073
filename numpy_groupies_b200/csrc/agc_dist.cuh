// agc_dist.cuh -- device helpers of the multi-GPU path (numpy_groupies_b200/distributed.py).
//   * pack / unpack: every partial table of a call (f64 sums, i64 counts / keys / indices, u8 flags) plus the status word
//     travels in ONE homogeneous buffer, so a merge is ONE all-reduce whatever the function (the round-1 merge issued one
//     collective per table and another one for the status);
//   * rank fold + carry apply: the cross-GPU carry exchange of the N->N functions (cumsum / cumprod / cummax / cummin):
//     every rank scans its own shard, the per-group totals of all ranks are all-gathered, rank r folds the totals of the
//     ranks before it (in rank order: deterministic) and adds that carry to its local scan in a second light pass.
#pragma once
#include "agc_generic.cuh"

namespace agc {

enum { PK_COPY = 0, PK_NEG = 1, PK_FLAG = 2, PK_NEGFLAG = 3 };   // how a table rides in the carrier (see aggcuda.h)
constexpr int PK_MAX_SEGS = 8;
struct PackSeg { void* table; long long count; int dtype; int xform; };
struct PackParams { PackSeg seg[PK_MAX_SEGS]; long long start[PK_MAX_SEGS + 1]; int nseg; int carrier; void* packed; int unpack; };

__device__ __forceinline__ long long pk_load_i(const void* t, int dt, long long i) {
  if (dt == AGC_DT_F64) return (long long)((const double*)t)[i];
  if (dt == AGC_DT_I32) return ((const int*)t)[i];
  if (dt == AGC_DT_U8 || dt == AGC_DT_BOOL) return ((const unsigned char*)t)[i];
  return ((const long long*)t)[i];
}
__global__ void k_pack_tables(const PackParams p) {
  const long long total = p.start[p.nseg];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
    int s = 0;
    while (s + 1 < p.nseg && e >= p.start[s + 1]) ++s;
    const PackSeg& sg = p.seg[s];
    const long long i = e - p.start[s];
    const bool neg = sg.xform == PK_NEG || sg.xform == PK_NEGFLAG;
    const bool flag = sg.xform == PK_FLAG || sg.xform == PK_NEGFLAG;
    if (!p.unpack) {
      if (p.carrier == AGC_DT_F64) {
        double v = sg.dtype == AGC_DT_F64 ? ((const double*)sg.table)[i] : (double)pk_load_i(sg.table, sg.dtype, i);
        ((double*)p.packed)[e] = neg ? -v : v;
      } else {
        long long v = pk_load_i(sg.table, sg.dtype, i);
        ((long long*)p.packed)[e] = neg ? ~v : v;               // bit complement: order-reversing and free of overflow
      }
    } else {
      if (p.carrier == AGC_DT_F64) {
        double v = ((const double*)p.packed)[e];
        if (neg) v = -v;
        if (flag) v = v != 0.0 ? 1.0 : 0.0;
        if (sg.dtype == AGC_DT_F64) ((double*)sg.table)[i] = v;
        else if (sg.dtype == AGC_DT_I32) ((int*)sg.table)[i] = (int)v;
        else if (sg.dtype == AGC_DT_U8 || sg.dtype == AGC_DT_BOOL) ((unsigned char*)sg.table)[i] = (unsigned char)v;
        else ((long long*)sg.table)[i] = (long long)v;
      } else {
        long long v = ((const long long*)p.packed)[e];
        if (neg) v = ~v;
        if (flag) v = v != 0;
        if (sg.dtype == AGC_DT_F64) ((double*)sg.table)[i] = (double)v;
        else if (sg.dtype == AGC_DT_I32) ((int*)sg.table)[i] = (int)v;
        else if (sg.dtype == AGC_DT_U8 || sg.dtype == AGC_DT_BOOL) ((unsigned char*)sg.table)[i] = (unsigned char)v;
        else ((long long*)sg.table)[i] = v;
      }
    }
  }
}

// carry[g] = fold over the ranks before `rank` of gathered[r][g] (T0 representation of the matching reduction:
// f64 / i64 sums and products, order-preserving int64 keys for max / min); op: 0 sum, 1 max, 2 min, 3 prod
__global__ void k_rank_fold(const void* gathered, int rank, long long size, int is_f64, int op, void* carry) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < size; g += stride) {
    if (is_f64) {
      const double* t = (const double*)gathered;
      double acc = op == 3 ? 1.0 : 0.0;
      for (int r = 0; r < rank; ++r) { const double v = t[(long long)r * size + g]; acc = op == 3 ? acc * v : acc + v; }
      ((double*)carry)[g] = acc;
    } else {
      const long long* t = (const long long*)gathered;
      long long acc = op == 0 ? 0 : (op == 3 ? 1 : (op == 1 ? AGC_EMPTY_MAX : AGC_EMPTY_MIN));
      for (int r = 0; r < rank; ++r) {
        const long long v = t[(long long)r * size + g];
        if (op == 0) acc = (long long)((unsigned long long)acc + (unsigned long long)v);
        else if (op == 3) acc = (long long)((unsigned long long)acc * (unsigned long long)v);
        else if (op == 1) acc = v > acc ? v : acc;
        else acc = v < acc ? v : acc;
      }
      ((long long*)carry)[g] = acc;
    }
  }
}

struct CarryParams {
  const void* labels; int ldtype; long long n, size;
  void* out; int out_dtype;      // the local scan, updated in place
  const void* carry; int acc_f64;  // T0 representation (see k_rank_fold)
  int op;                        // 0 sum, 1 max, 2 min, 3 prod
  int a_float, a_u64;
  int* status;
};
__device__ __forceinline__ double ld_out_f(const void* o, int dt, long long i) { return dt == AGC_DT_F32 ? (double)((const float*)o)[i] : ((const double*)o)[i]; }
__global__ void k_apply_carry(const CarryParams p) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n; i += stride) {
    const long long g = load_int(p.labels, p.ldtype, i);
    if (g < 0 || g >= p.size) continue;                     // the local scan already flagged it
    const bool out_fl = dtype_is_float(p.out_dtype);
    if (p.op == 0 || p.op == 3) {
      if (p.acc_f64) {
        const double c = ((const double*)p.carry)[g];
        if (out_fl) { const double v = ld_out_f(p.out, p.out_dtype, i); store_from_f64(p.out, p.out_dtype, i, p.op == 0 ? c + v : c * v); }
        else { const double v = (double)load_int(p.out, p.out_dtype, i); store_from_f64(p.out, p.out_dtype, i, p.op == 0 ? c + v : c * v); }
      } else {
        const unsigned long long c = (unsigned long long)((const long long*)p.carry)[g];
        if (out_fl) { const double v = ld_out_f(p.out, p.out_dtype, i); store_from_f64(p.out, p.out_dtype, i, p.op == 0 ? (double)(long long)c + v : (double)(long long)c * v); }
        else {
          const unsigned long long v = (unsigned long long)load_int(p.out, p.out_dtype, i);
          store_from_i64(p.out, p.out_dtype, i, (long long)(p.op == 0 ? c + v : c * v), p.a_u64);
        }
      }
    } else {
      const long long ck = ((const long long*)p.carry)[g];
      const long long ident = p.op == 1 ? AGC_EMPTY_MAX : AGC_EMPTY_MIN;
      if (ck == ident) continue;                            // no earlier rank saw this group
      if (p.a_float) {
        const double v = ld_out_f(p.out, p.out_dtype, i);
        const long long nankey = p.op == 1 ? AGC_EMPTY_MIN : AGC_EMPTY_MAX;
        if (ck == nankey) { store_from_f64(p.out, p.out_dtype, i, __longlong_as_double(0x7ff8000000000000ll)); continue; }
        const double c = unkey_f64(ck);
        if (v != v) continue;                               // a NaN of the local scan stays
        store_from_f64(p.out, p.out_dtype, i, p.op == 1 ? (c > v ? c : v) : (c < v ? c : v));
      } else {
        const long long raw = load_int(p.out, p.out_dtype, i);
        const long long vk = p.a_u64 ? key_of((unsigned long long)raw) : raw;
        const long long rk = p.op == 1 ? (ck > vk ? ck : vk) : (ck < vk ? ck : vk);
        store_from_i64(p.out, p.out_dtype, i, p.a_u64 ? (long long)((unsigned long long)rk ^ 0x8000000000000000ull) : rk, p.a_u64);
      }
    }
  }
}

// first / last across shards: after the index merge every rank knows the winning GLOBAL element index of a group, but only
// the rank that holds that element can read its value.  pack: carrier[g] = value bits where this shard owns the winner,
// else 0; (int64 SUM all-reduce: at most one rank contributes per group); unpack: non-empty groups take the carrier's bits.
__global__ void k_owner_pack(void* out, int elem_bytes, long long size, const long long* t1, long long index_base, long long n,
                             long long* carrier, int unpack) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < size; g += stride) {
    const long long idx = t1[g];
    const bool empty = idx == AGC_NO_INDEX || idx < 0;
    if (!unpack) {
      const long long loc = idx - index_base;
      const bool mine = !empty && loc >= 0 && loc < n;
      unsigned long long bits = 0;
      if (mine) {
        if (elem_bytes == 8) bits = ((const unsigned long long*)out)[g];
        else if (elem_bytes == 4) bits = ((const unsigned*)out)[g];
        else if (elem_bytes == 2) bits = ((const unsigned short*)out)[g];
        else bits = ((const unsigned char*)out)[g];
      }
      carrier[g] = (long long)bits;
    } else if (!empty) {
      const unsigned long long bits = (unsigned long long)carrier[g];
      if (elem_bytes == 8) ((unsigned long long*)out)[g] = bits;
      else if (elem_bytes == 4) ((unsigned*)out)[g] = (unsigned)bits;
      else if (elem_bytes == 2) ((unsigned short*)out)[g] = (unsigned short)bits;
      else ((unsigned char*)out)[g] = (unsigned char)bits;
    }
  }
}

}  // namespace agc

// ---------------------------------------------------------------------------------------------
// deterministic fixed-order mode (north_star: "optional deterministic fixed-order mode")
// ---------------------------------------------------------------------------------------------
// Floating sums / means / variances / products whose bits do not depend on how the hardware schedules atomics: the
// elements are first ordered by group with the stable radix sort (agc_group_order: order[] + CSR offsets[]), then ONE WARP
// per group walks its elements in their original order -- lane l takes elements l, l + 32, ... of the group and adds them
// one after the other in f64, the 32 lane totals are combined by a fixed butterfly.  The order of every addition is a
// function of the group's element list alone: same input, same bits, on every run and on every B200.  var / std run the
// reference's second pass (aggregate_numpy.py:180-187) the same way.  Results land in the ordinary accumulator tables and
// are finalised by k_finalize like any other call.
namespace agc {

struct OrderedParams {
  const void* a; int a_dtype;
  const long long* order; const long long* offsets;
  long long size;
  Tables t;
  int op, flags;
};

__device__ __forceinline__ double ord_load(const void* a, int dt, long long i) {
  return dt == AGC_DT_F32 ? (double)((const float*)a)[i] : ((const double*)a)[i];
}
__device__ __forceinline__ double warp_fold_add(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_fold_mul(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) k_ordered_accumulate(const OrderedParams p) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool nanskip = p.flags & AGC_F_NANSKIP;
  const bool pass2_only = p.flags & AGC_F_PASS2;
  for (long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < p.size; g += warps) {
    const long long lo = p.offsets[g], hi = p.offsets[g + 1];
    double acc = p.op == AGC_OP_PROD ? 1.0 : 0.0;
    long long cnt = 0;
    if (!pass2_only) {
      for (long long k = lo + lane; k < hi; k += 32) {
        const double v = ord_load(p.a, p.a_dtype, p.order[k]);
        if (nanskip && v != v) continue;
        ++cnt;
        if (p.op == AGC_OP_PROD) acc *= v;
        else if (p.op == AGC_OP_SUMSQ) { const double s = p.a_dtype == AGC_DT_F32 ? (double)((float)v * (float)v) : v * v; acc += s; }   // square in a.dtype
        else acc += v;
      }
      acc = p.op == AGC_OP_PROD ? warp_fold_mul(acc) : warp_fold_add(acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (lane == 0) { p.t.t0f[g] = acc; p.t.t1[g] = cnt; p.t.b1[g] = cnt > 0; }
    } else {
      acc = p.t.t0f[g]; cnt = p.t.t1[g];                      // the sums and counts of ALL ranks (merged in rank order by the caller)
    }
    if (p.op == AGC_OP_VAR) {
      const double mean = acc / (double)cnt;
      double m2 = 0.0;
      for (long long k = lo + lane; k < hi; k += 32) {
        const double v = ord_load(p.a, p.a_dtype, p.order[k]);
        if (nanskip && v != v) continue;
        const double d = v - mean;
        m2 += d * d;
      }
      m2 = warp_fold_add(m2);
      if (lane == 0) p.t.t2[g] = m2;
    }
  }
}

}  // namespace agc
