/* aggcuda.h -- C ABI of libaggcuda.so, the sm_100a device side of `aggregate_cuda`.
 *
 * This is the drop-in boundary for the reference's per-function kernel layer:
 *   reference contract    f(group_idx, a, size, fill_value, dtype=None, **kw) -> ndarray
 *                         looked up in `_impl_dict` (numpy_groupies/aggregate_numpy.py:277-305),
 *                         driven by `_aggregate_base` (aggregate_numpy.py:308-367);
 *   numba counterpart     AggregateOp.__call__ / jitted `loop` (aggregate_numba.py:45-96,142-157).
 * One agc_aggregate() call replaces one `_impl_dict[func](...)` call *including* the label
 * flattening of Forms 3/4 (utils.py:386-432, 524-535), which is folded into the index-load
 * stage instead of being materialised.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer except `req` itself is a DEVICE pointer;
 *   - the library never allocates: outputs, workspace and the status words are caller-owned
 *     (torch tensors on the Python side); agc_workspace_bytes() sizes the workspace;
 *   - all work is enqueued on the caller's stream (`cudaStream_t` passed as void*), no hidden
 *     synchronisation; the caller reads `status` after it synchronises;
 *   - argument errors: negative return code + agc_last_error(); DATA errors (negative label,
 *     label >= size, Form-4 coordinate out of range -- the reference's ValueError cases,
 *     utils.py:458-459,521-522,533-534 / aggregate_numba.py:151-154) set status[0] != 0.
 */
#ifndef AGGCUDA_H
#define AGGCUDA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGC_ABI_VERSION 5

/* element dtypes (values, labels and outputs) */
enum {
  AGC_DT_BOOL = 0, AGC_DT_I8 = 1, AGC_DT_U8 = 2, AGC_DT_I16 = 3, AGC_DT_U16 = 4, AGC_DT_I32 = 5,
  AGC_DT_U32 = 6, AGC_DT_I64 = 7, AGC_DT_U64 = 8, AGC_DT_F32 = 9, AGC_DT_F64 = 10, AGC_DT_COUNT = 11
};

/* operations: one per reference kernel (aggregate_numpy.py line of the function it replaces) */
enum {
  AGC_OP_SUM = 0,      /* _sum            :20-39   */
  AGC_OP_PROD = 1,     /* _prod           :42-48   */
  AGC_OP_LEN = 2,      /* _len            :51-52   */
  AGC_OP_LAST = 3,     /* _last           :55-62   */
  AGC_OP_FIRST = 4,    /* _first          :65-69   */
  AGC_OP_ALL = 5,      /* _all            :72-78   */
  AGC_OP_ANY = 6,      /* _any            :81-87   */
  AGC_OP_MIN = 7,      /* _min            :90-99   */
  AGC_OP_MAX = 8,      /* _max            :102-111 */
  AGC_OP_ARGMAX = 9,   /* _argmax         :114-125 */
  AGC_OP_ARGMIN = 10,  /* _argmin         :128-139 */
  AGC_OP_MEAN = 11,    /* _mean           :142-161 */
  AGC_OP_SUMSQ = 12,   /* _sum_of_squres  :164-172 */
  AGC_OP_VAR = 13,     /* _var / _std     :175-199 (AGC_F_SQRT selects std) */
  AGC_OP_ALLNAN = 14,  /* _allnan         :202-203 */
  AGC_OP_ANYNAN = 15,  /* _anynan         :206-207 */
  /* N -> N operations (output has n elements) */
  AGC_OP_CUMSUM = 16,  /* _cumsum :244-266; per-group running sum in element order */
  AGC_OP_CUMPROD = 17, /* aggregate_numba.py CumProd :491-492 */
  AGC_OP_CUMMAX = 18,  /* aggregate_numba.py CumMax  :495-496 */
  AGC_OP_CUMMIN = 19,  /* aggregate_numba.py CumMin  :499-500 */
  AGC_OP_SORT = 20,    /* _sort :210-214 (AGC_F_REVERSE for reverse=True) */
  /* fused multi-statistic accumulation (README.md:185 "min+max in one pass"): only with AGC_F_ACC_ONLY; leaves max keys in
   * T0 and min keys in T2, which AGC_F_FIN_ONLY calls with op = MAX / op = MIN | AGC_F_KEYS_IN_T2 turn into results */
  AGC_OP_MINMAX = 21,
  AGC_OP_COUNT = 22
};

/* request flags */
#define AGC_F_NANSKIP        0x0001  /* nan-variant: NaN elements do not exist (aggregate_numpy.py:336-349)   */
#define AGC_F_SQRT           0x0002  /* AGC_OP_VAR -> std                                                    */
#define AGC_F_REVERSE        0x0004  /* AGC_OP_SORT descending                                               */
#define AGC_F_ASSUME_SORTED  0x0008  /* caller promises non-decreasing flat labels (verified on device)       */
#define AGC_F_NO_FAST        0x0010  /* force the general global-table kernels (testing / A-B measurements)  */
#define AGC_F_NEED_TOUCHED   0x0020  /* fill_value differs from the op's identity: track never-touched groups */
#define AGC_F_ACC_ONLY       0x0040  /* stop after accumulation: partial tables stay in the workspace         */
#define AGC_F_FIN_ONLY       0x0080  /* only finalise partial tables already in the workspace (multi-GPU)     */
#define AGC_F_FILL_NAN       0x0100  /* fill_value is NaN (mean/var keep the raw x/0, aggregate_numpy.py:154) */
#define AGC_F_NO_INIT        0x0200  /* with ACC_ONLY: do not re-initialise tables (accumulate more shards)   */
#define AGC_F_PASS2          0x0400  /* with ACC_ONLY: run only the second pass of var/std/arg* over tables   */
                                     /* the caller has merged (multi-GPU: sum,count / extreme key)            */
#define AGC_F_TOUCHED_BY_COUNT 0x0800 /* FIN_ONLY of sum over tables a mean / var pass left: touched = count > 0 */
#define AGC_F_KEYS_IN_T2     0x1000  /* FIN_ONLY of min after AGC_OP_MINMAX: the min keys live in T2            */
#define AGC_F_DETERMINISTIC  0x2000  /* float sums in a fixed order: same bits for any kernel schedule / GPU count */

/* how the flat group of element i is obtained (input_validation's three shapes, utils.py:435-542) */
enum {
  AGC_FORM_FLAT = 0,   /* Form 1/2: labels[n]                                                  */
  AGC_FORM_AXIS = 1,   /* Form 3:   labels[a.shape[axis]], a is ndim-dimensional, C order      */
  AGC_FORM_MULTI = 2   /* Form 4/5: labels[ndim][n], flat = sum_d labels[d][i] * strides[d]     */
};
#define AGC_MAX_NDIM 8

typedef struct agc_index {
  int32_t form;                    /* AGC_FORM_*                                                       */
  int32_t dtype;                   /* label dtype (any integer AGC_DT_*)                               */
  const void* labels;              /* device pointer, see form                                         */
  int32_t ndim;                    /* AXIS: a.ndim; MULTI: number of label rows                        */
  int32_t axis;                    /* AXIS: the reduced axis                                           */
  int64_t dims[AGC_MAX_NDIM];      /* AXIS: a.shape; MULTI: size per coordinate (bounds check)         */
  int64_t strides[AGC_MAX_NDIM];   /* AXIS: output stride per a-dimension (the axis entry multiplies   */
                                   /* the label); MULTI: stride per coordinate                         */
  int64_t axis_size;               /* AXIS: number of groups along the axis (bounds check)             */
} agc_index_t;

typedef struct agc_request {
  int32_t op;                      /* AGC_OP_*                                                          */
  int32_t flags;                   /* AGC_F_*                                                           */
  int32_t a_dtype;                 /* dtype of `a` (or of the scalar when a == NULL)                    */
  int32_t out_dtype;               /* dtype written to `out`                                            */
  const void* a;                   /* device, n elements, C-contiguous; NULL => scalar a (Form 2/5)     */
  double a_scalar_f64;             /* scalar a when a == NULL and a_dtype is floating                   */
  int64_t a_scalar_i64;            /* scalar a when a == NULL and a_dtype is integral                   */
  agc_index_t index;
  int64_t n;                       /* number of elements                                                */
  int64_t size;                    /* number of flat groups                                             */
  void* out;                       /* device: `size` elements (reductions) or `n` elements (N->N ops)   */
  double fill_f64;                 /* fill_value for floating outputs                                   */
  int64_t fill_i64;                /* fill_value for integral / bool outputs                            */
  double ddof;                     /* AGC_OP_VAR                                                        */
  int64_t index_base;              /* global index of element 0 (shards: first/last/arg*)               */
  int64_t arg_stride;              /* arg* on Form 3: report (i / arg_stride) % arg_len ...             */
  int64_t arg_len;                 /* ... 0 => report the flat element index                            */
  void* workspace;                 /* device scratch, >= agc_workspace_bytes(req)                       */
  int64_t workspace_bytes;
  int32_t* status;                 /* device int32[8]; [0] bad label, [1] fast-path precision guard,    */
                                   /* [2] labels-not-sorted (ASSUME_SORTED violated), [3] regime used,  */
                                   /* [4..6] label probe: skewed?, largest bucket, equal neighbours;    */
                                   /* [7] library scratch (Form 3 work counter / probe: values span > window) */
} agc_request_t;

/* one partial table inside the workspace after AGC_F_ACC_ONLY (for the multi-GPU merge) */
typedef struct agc_partial {
  int64_t offset_bytes;            /* from the workspace base                                           */
  int64_t count;                   /* elements                                                          */
  int32_t dtype;                   /* AGC_DT_F64 / AGC_DT_I64 / AGC_DT_U8                               */
  int32_t reduce;                  /* 0 = sum, 1 = max, 2 = min, 3 = prod                               */
} agc_partial_t;

int         agc_abi_version(void);
const char* agc_last_error(void);
/* bytes of device scratch agc_aggregate(req) needs (depends on op, n, size). */
int64_t     agc_workspace_bytes(const agc_request_t* req);
/* describe the partial tables agc_aggregate(req | AGC_F_ACC_ONLY) leaves behind; returns their number */
int         agc_partials(const agc_request_t* req, agc_partial_t* out, int max_out);
/* run one aggregation on `stream`; 0 on success */
int         agc_aggregate(const agc_request_t* req, void* stream);
/* labels.min()/max() into out_minmax[2] (device int64) -- size=None support (utils.py:515-526) */
int         agc_label_minmax(const void* labels, int32_t dtype, int64_t n, int64_t* out_minmax, void* stream);
/* ret[group_idx] gather (`unpack`, utils.py:548-553); elem_bytes in {1,2,4,8} */
int         agc_unpack(const void* table, int32_t elem_bytes, int64_t size, const void* labels, int32_t dtype,
                       int64_t n, void* out, int32_t* status, void* stream);
/* stable sort of (label, position) pairs by label: fills order[n] (int64 positions) and, if non-NULL,
 * offsets[size+1] (CSR group starts) -- the device half of `_array` / generic callables
 * (aggregate_numpy.py:217-241). */
int64_t     agc_group_order_workspace_bytes(int64_t n, int64_t size);
int         agc_group_order(const agc_index_t* index, int64_t n, int64_t size, int64_t* order, int64_t* offsets,
                            void* workspace, int64_t workspace_bytes, int32_t* status, void* stream);

/* ---- label utilities (the steps either side of aggregate): one chained single-pass scan over int64 ----
 * mode 0 step_indices (aggregate_numba.py:591-605): out = int64[total + 1] run starts of x followed by n
 *      1 step_count   (aggregate_numba.py:577-588): only *total = number of runs
 *      2 label_contiguous_1d (utils.py:597-632): out = int64[n]; x bool mask or integer values
 *      3 relabel table of relabel_groups_masked (utils.py:651-681): x = keep flags (size entries), out[k] = new label of k
 *        in integer dtype out_dtype
 *      4 inclusive running sum of the integers x into int64 out (multi_arange, utils.py:572-594)
 * `total` (device int64, may be NULL) receives the sum of all per-element flags / values. */
int64_t     agc_label_scan_workspace_bytes(int64_t n);
int         agc_label_scan(int32_t mode, const void* x, int32_t x_dtype, int64_t n, void* out, int32_t out_dtype,
                           int64_t* total, void* workspace, int64_t workspace_bytes, void* stream);
/* keep[g] = 1 for every label g that occurs (keep is uint8[size], zeroed by the caller): relabel_groups_unique, utils.py:635-648 */
int         agc_mark_present(const void* labels, int32_t dtype, int64_t n, int64_t size, void* keep, int32_t* status, void* stream);
/* multi_arange (utils.py:572-594): incl = inclusive running sums of n (int64[m]), out = int64[total] */
int         agc_multi_arange(const int64_t* incl, int64_t m, int64_t total, int64_t* out, void* stream);

/* ---- multi-GPU helpers (numpy_groupies_b200/distributed.py) -------------------------------------------------
 * agc_pack_tables: gather up to 8 tables of different dtypes (float64 / int64 / int32 / uint8) into ONE homogeneous
 * buffer (`carrier_dtype` float64 or int64) so that a merge is a single all-reduce, or (unpack != 0) scatter the reduced
 * buffer back.  xform per table: 0 value as is, 1 negated (a MAX table riding in a MIN reduction and vice versa),
 * 2 flag (a 0/1 table riding in a SUM: any non-zero reads back as 1), 3 negated flag. */
typedef struct agc_pack_seg { void* table; int64_t count; int32_t dtype; int32_t xform; } agc_pack_seg_t;
int         agc_pack_tables(const agc_pack_seg_t* segs, int32_t nseg, void* packed, int32_t carrier_dtype, int32_t unpack, void* stream);
/* cross-GPU carry exchange of the N->N functions: gathered = [world][size] per-group totals of every rank in the T0
 * representation of the matching reduction (float64 / int64 sums and products, int64 order-preserving keys for max / min;
 * reduce: 0 sum, 1 max, 2 min, 3 prod).  carry[g] = fold of the ranks before `rank`, in rank order. */
int         agc_rank_fold(const void* gathered, int32_t rank, int64_t size, int32_t dtype, int32_t reduce, void* carry, void* stream);
/* out[i] = carry[labels[i]] (op) out[i] for op in AGC_OP_CUMSUM / CUMPROD / CUMMAX / CUMMIN (out = this rank's local scan) */
int         agc_apply_carry(int32_t op, const void* labels, int32_t label_dtype, int64_t n, int64_t size, void* out, int32_t out_dtype,
                            int32_t a_dtype, const void* carry, int32_t carry_dtype, int32_t* status, void* stream);

/* deterministic fixed-order mode: instead of agc_aggregate(req | AGC_F_ACC_ONLY), fill the accumulator tables of a floating
 * sum / sumofsquares / mean / var(std) / prod by walking every group's elements in their original order (`order`, `offsets`
 * from agc_group_order): one warp per group, lane-strided sequential f64 additions, fixed butterfly across the lanes.  The
 * result's bits depend on the input alone.  With AGC_F_PASS2 only the second pass of var / std runs (T0 / T1 hold the sums and
 * counts merged over the shards; T2 receives this shard's squared deviations from the global mean).  Finalise with
 * agc_aggregate(req | AGC_F_FIN_ONLY). */
int         agc_ordered_accumulate(const agc_request_t* req, const int64_t* order, const int64_t* offsets, void* stream);

/* first / last over shards (after the index tables were merged and req was finalised with AGC_F_FIN_ONLY): pack writes
 * carrier[g] = bits of out[g] where THIS shard holds the group's winning element, else 0 (one int64 SUM all-reduce hands
 * every rank the owner's bits); unpack stores them into out[g] for every non-empty group. */
int         agc_owner_pack(const agc_request_t* req, int64_t* carrier, int32_t unpack, void* stream);

/* float16 values / results (check_dtype keeps float16 for float16 input, utils.py:435-470; the kernels compute them as
 * float32): n elements float16 -> float32 (to_half == 0, exact) or float32 -> float16 (round to nearest even, as
 * numpy's astype), device pointers, src != dst. */
int         agc_convert_f16(const void* src, void* dst, int64_t n, int32_t to_half, void* stream);

/* Host side of the H2D path for numpy labels (the reference hands `group_idx` to its kernels as it is,
 * aggregate_numpy.py:370-392; here the labels have to cross PCIe first, and for int64 labels that is 8 of the 12 bytes
 * per element): narrows n 8-byte integer labels (AGC_DT_I64 / AGC_DT_U64) to int32 into dst -- typically a page-locked
 * staging buffer -- mapping every label outside [0, 2^31) to -1, which the kernels report as a bad label exactly like
 * the original value.  Pure host code, thread-safe, no CUDA call; callers run it on slices from several threads. */
int         agc_host_narrow_labels(const void* src, int32_t src_dtype, int64_t n, int32_t* dst);
/* The same staging step for everything else: a copy with non-temporal stores (the destination is read next by the DMA engine,
 * not by this core: no read-for-ownership, no cache pollution). */
int         agc_host_copy_stream(const void* src, void* dst, int64_t bytes);

/* measurement hooks (bench.py): when enabled, agc_aggregate brackets its dominant accumulate kernel with
 * CUDA events on the caller's stream -- a ring of 64 pairs, one per call, so a timed loop needs no read-back;
 * agc_profile_ms(back) waits for the call `back` calls ago (0 = the last one) and returns its duration,
 * agc_profile_last_ms() = agc_profile_ms(0).
 * agc_launch_count() = number of libaggcuda kernels launched by this process so far. */
int         agc_profile_enable(int on);
float       agc_profile_last_ms(void);
float       agc_profile_ms(int back);
int64_t     agc_launch_count(void);
/* kernel-selection knobs for A/B measurements and tests; value < 0 only reads.  Returns the (new) value.
 *   key 0  strategy for tables larger than one SM's shared memory: 0 hot-key cache, 1 on-device label probe
 *          decides (default), 3 always partial table + global RED, 4 always the run kernel (segmented warp reduce)
 *   key 1  CTAs per SM of the hot-key-cache kernel (1 | 2)
 *   key 2  large tables: one-word sums with carry forwarding and split mean (1, default) or fused tables (0)
 *   key 3  bytes of shared memory one CTA per SM may use for its table (default 199552: <= 196 KB per SM keeps 60 KB
 *          of L1 for the global loads in flight; the 228 KB carve-out costs a third of the streaming bandwidth)
 *   key 4  the same per CTA for kernels that run two CTAs per SM (default 99200)
 *   key 5  a partial table keeps the key-3 size while at most this percentage of the groups stays uncovered; beyond
 *          that it grows to the full 226 KB (default 10)
 *   key 6  sorted-label scans: 1 single-pass chained scan with warp-parallel look-back (default), 0 the three-kernel
 *          reduce / scan-tiles / apply variant
 *   key 7  sums / means on tables larger than one SM: thread-block clusters of this many SMs (2 | 4) own the table together
 *          and exchange records over DSMEM (k_team); 0 = partial table + global RED (default: the exchange measured slower than the REDs it saves)
 *   key 8  k_team runs while at most this percentage of the groups stays without a slot (default 60)
 *   key 9  CTAs of the partial-table kernel (0 = one per SM, default): the L2 takes scattered REDs best from 111-129 senders */
int         agc_tuning(int key, int value);

#ifdef __cplusplus
}
#endif
#endif /* AGGCUDA_H */
