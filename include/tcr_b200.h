/*
 * tcr_b200.h — C-ABI of the B200 evaluation back end for TEQ functor graphs.
 *
 * This is the drop-in boundary of the hot path: every entry point below is what a
 * `cuda::` data holder (the replacement for the reference's eigen::TensOp / MatOp /
 * TensAssign, internal/eigen/device.hpp:295-544) binds when
 * `egen::typed_exec<T>(opcode, out, shape, args, attrs)` (generated from
 * cfg/ops.yml:6,63-711 by tools/egen/plugins/opcodes.py:36-46; sole caller
 * tenncor/eteq/functor.hpp:201-202) is re-pointed at the device.
 *
 * Conventions
 *  - plain pointers and sizes only; all data pointers are DEVICE pointers unless the
 *    parameter name starts with `host`;
 *  - tensors are column-major rank-8 (teq dim 0 fastest, internal/eigen/convert.hpp:33,
 *    internal/teq/shape.hpp:56-59); shapes are `int64_t[8]`, unused ranks = 1;
 *  - dtype codes are the reference's generated `egen::_GENERATED_DTYPE` values
 *    (cfg/fulltype.yml:4-34, tools/egen/plugins/dtypes.py:20-27);
 *  - every function returns 0 on success, non-zero on failure with the reason in
 *    `tcr_last_error()`; the C++ host converts that to `global::fatal`
 *    (internal/global/logs.hpp:35-38,70-81). There is NO CPU fallback: when no CUDA
 *    device is present every compute entry point fails with TCR_ERR_NODEVICE;
 *  - all work is enqueued on the library stream (`tcr_stream()`), nothing synchronises
 *    unless documented.
 */
#ifndef TCR_B200_H
#define TCR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCR_RANK_CAP 8 /* teq::rank_cap, internal/teq/shape.hpp:45 */

/* status codes */
enum {
  TCR_OK = 0,
  TCR_ERR_CUDA = 1,     /* CUDA runtime error (text in tcr_last_error) */
  TCR_ERR_ARG = 2,      /* invalid argument */
  TCR_ERR_DTYPE = 3,    /* dtype not supported by this kernel */
  TCR_ERR_NODEVICE = 4, /* no CUDA device / tcr_init not called */
  TCR_ERR_NCCL = 5,
  TCR_ERR_UNSUPPORTED = 6
};

/* egen::_GENERATED_DTYPE (cfg/fulltype.yml order; BAD_TYPE = 0) */
enum {
  TCR_BAD_TYPE = 0,
  TCR_DOUBLE = 1,
  TCR_FLOAT = 2,
  TCR_INT8 = 3,
  TCR_UINT8 = 4,
  TCR_INT16 = 5,
  TCR_UINT16 = 6,
  TCR_INT32 = 7,
  TCR_UINT32 = 8,
  TCR_INT64 = 9,
  TCR_UINT64 = 10
};

/* egen::_GENERATED_OPCODE (cfg/ops.yml:63-711 order; BAD_OP = 0) */
enum {
  TCR_OP_BAD = 0,
  TCR_OP_IDENTITY = 1, TCR_OP_ABS, TCR_OP_NEG, TCR_OP_SIN, TCR_OP_COS, TCR_OP_TAN,
  TCR_OP_EXP, TCR_OP_LOG, TCR_OP_SQRT, TCR_OP_ROUND, TCR_OP_SIGMOID, TCR_OP_TANH,
  TCR_OP_SQUARE, TCR_OP_CUBE, TCR_OP_RAND_UNIF, TCR_OP_REVERSE, TCR_OP_REDUCE_SUM,
  TCR_OP_REDUCE_PROD, TCR_OP_REDUCE_MIN, TCR_OP_REDUCE_MAX, TCR_OP_ARGMAX,
  TCR_OP_PERMUTE, TCR_OP_EXTEND, TCR_OP_RESHAPE, TCR_OP_SLICE, TCR_OP_PAD,
  TCR_OP_STRIDE, TCR_OP_SCATTER, TCR_OP_POW, TCR_OP_ADD, TCR_OP_SUB, TCR_OP_MUL,
  TCR_OP_DIV, TCR_OP_MIN, TCR_OP_MAX, TCR_OP_EQ, TCR_OP_NEQ, TCR_OP_LT, TCR_OP_GT,
  TCR_OP_MATMUL, TCR_OP_CONTRACT, TCR_OP_CONV, TCR_OP_SELECT, TCR_OP_CONCAT,
  TCR_OP_ASSIGN, TCR_OP_ASSIGN_ADD, TCR_OP_ASSIGN_SUB, TCR_OP_ASSIGN_MUL,
  TCR_OP_ASSIGN_DIV, TCR_OP_CAST, /* = 50 */
  TCR_OP_COUNT
};

/* ------------------------------------------------------------------ runtime */

/* Select `device`, create the library stream, query SM count. Idempotent. */
int tcr_init(int device);
int tcr_shutdown(void);
/* Thread-local text of the last failure ("" if none). */
const char* tcr_last_error(void);
int tcr_device_count(void); /* 0 when no driver / no GPU; never fails */
int tcr_sm_count(void);
void* tcr_stream(void); /* cudaStream_t of the library */
int tcr_sync(void);     /* cudaStreamSynchronize(tcr_stream()) */

/* Device arena backing cuda::DeviceRuntimeMemory (replaces eigen::RuntimeMemory's
 * malloc/free, internal/eigen/memory.hpp:26-37). Stream-ordered, size-bucketed
 * free lists: a freed block is reusable by later work on the library stream. */
int tcr_alloc(void** out, size_t bytes);
int tcr_free(void* ptr);
int tcr_arena_stats(size_t* bytes_in_use, size_t* bytes_reserved, size_t* n_device_mallocs);
int tcr_arena_trim(void); /* return cached blocks to the driver */

/* Copies (async on the library stream; host buffers should be pinned for overlap). */
int tcr_host_alloc(void** out, size_t bytes); /* pinned */
int tcr_host_free(void* ptr);
int tcr_h2d(void* dst, const void* host_src, size_t bytes);
/* Input prefetch (extension; the reference's Variable::assign is a synchronous memcpy,
 * tenncor/eteq/variable.hpp:55-90): `tcr_h2d_prefetch` copies a pinned host batch into a device
 * staging buffer on a dedicated copy stream, so it overlaps the step that is computing;
 * `tcr_prefetch_commit` makes the library stream wait for that copy and moves staging -> dst
 * (HBM -> HBM). A later prefetch into the same staging buffer waits for the commit that read it. */
int tcr_h2d_prefetch(void* staging, const void* host_src, size_t bytes);
int tcr_prefetch_commit(void* dst, const void* staging, size_t bytes);
int tcr_prefetch_sync(void); /* wait for the copy stream (tcr_sync covers the library stream only) */
int tcr_d2h(void* host_dst, const void* src, size_t bytes); /* async; tcr_sync() before reading */
int tcr_d2d(void* dst, const void* src, size_t bytes);
int tcr_memset(void* dst, int byte, size_t bytes);

/* Events for device-side timing on the library stream. */
int tcr_event_create(void** out);
int tcr_event_destroy(void* ev);
int tcr_event_record(void* ev);
int tcr_event_sync(void* ev); /* wait for the work queued before the record (a deferred tcr_d2h: read the host buffer after this) */
int tcr_event_elapsed_ms(void* start, void* stop, float* ms); /* synchronises on stop */

/* CUDA-graph capture of a launch sequence (replaces the per-node host traversal of
 * teq::TravEvaluator::visit_func, internal/teq/evaluator.hpp:34-43, on replay). */
int tcr_graph_begin(void);
/* Capture lanes. Between tcr_graph_begin and tcr_graph_end the launcher may route launches to
 * side lanes (lane 0 = the library stream) and order lanes with marks; independent steps then
 * become parallel branches of the instantiated graph. tcr_graph_end joins every lane. Scratch
 * memory obtained with tcr_alloc during capture is recycled per lane only. */
#define TCR_GRAPH_LANES 8
int tcr_graph_lane(int lane);
int tcr_graph_record(int* mark_id);
int tcr_graph_wait(int mark_id);
int tcr_graph_end(void** out_exec);
int tcr_graph_launch(void* exec);
int tcr_graph_destroy(void* exec);

/* Number of kernels launched by this library since init (for bench "gpu_launches"). */
uint64_t tcr_launch_count(void);

/* --------------------------------------------------------------- elementwise */

/* Fused elementwise program: a tiny register machine evaluated per element.
 * Replaces one or a chain of the cwise factories of internal/eigen/operator.hpp:377-987
 * (+ select :1050-1067, cast :1239-1260, assign* :1190-1237).
 *
 * Registers r0..r7 hold values of the compute type of the program's `dtype`.
 * Input i is loaded into register i before the first instruction. Each
 * instruction is `dst = op(a, b, c)`; TCR_EW_CONST loads `imm`. Outputs are stored
 * from their registers after the last instruction. */
enum {
  TCR_EW_NOP = 0,
  /* unary: opcode values equal the reference opcode for ABS..CUBE */
  TCR_EW_ABS = TCR_OP_ABS, TCR_EW_NEG = TCR_OP_NEG, TCR_EW_SIN = TCR_OP_SIN,
  TCR_EW_COS = TCR_OP_COS, TCR_EW_TAN = TCR_OP_TAN, TCR_EW_EXP = TCR_OP_EXP,
  TCR_EW_LOG = TCR_OP_LOG, TCR_EW_SQRT = TCR_OP_SQRT, TCR_EW_ROUND = TCR_OP_ROUND,
  TCR_EW_SIGMOID = TCR_OP_SIGMOID, TCR_EW_TANH = TCR_OP_TANH,
  TCR_EW_SQUARE = TCR_OP_SQUARE, TCR_EW_CUBE = TCR_OP_CUBE,
  /* binary */
  TCR_EW_POW = TCR_OP_POW, TCR_EW_ADD = TCR_OP_ADD, TCR_EW_SUB = TCR_OP_SUB,
  TCR_EW_MUL = TCR_OP_MUL, TCR_EW_DIV = TCR_OP_DIV, TCR_EW_MIN = TCR_OP_MIN,
  TCR_EW_MAX = TCR_OP_MAX, TCR_EW_EQ = TCR_OP_EQ, TCR_EW_NEQ = TCR_OP_NEQ,
  TCR_EW_LT = TCR_OP_LT, TCR_EW_GT = TCR_OP_GT,
  /* ternary: dst = a != 0 ? b : c */
  TCR_EW_SELECT = TCR_OP_SELECT,
  /* machine ops */
  TCR_EW_MOV = 64,   /* dst = a */
  TCR_EW_CONST = 65  /* dst = imm */
};

#define TCR_EW_MAX_INPUTS 8
#define TCR_EW_MAX_OUTPUTS 4
#define TCR_EW_MAX_INSTRS 32
#define TCR_EW_NREGS 8

typedef struct {
  uint8_t op, dst, a, b, c;
  uint8_t _pad[3];
  double imm;
} tcr_ew_instr;

/* An input is either the full iteration space or a broadcast of it. The output index
 * i (column-major linear) is split as i = i0 + D0*(i1 + D1*i2) with D = prog.dims; the
 * input holds extents e[k] in {1, D[k]} and is read at
 * j = (e0>1?i0:0) + e0*((e1>1?i1:0) + e1*(e2>1?i2:0)).
 * This covers EXTEND of a scalar, of a leading block of dims (bias [H] -> [H,B]) and
 * of a trailing block (reduce gradient [1,B] -> [H,B]) without materialising the
 * broadcast (reference materialises it: internal/eigen/operator.hpp:159-173). */
typedef struct {
  const void* ptr;
  int32_t dtype;      /* element type in memory; converted to the compute type on load */
  uint8_t bcast[3];   /* bcast[k] = 1 -> extent 1 along segment k */
  uint8_t _pad;
} tcr_ew_input;

typedef struct {
  void* ptr;
  int32_t dtype;
  uint8_t reg;
  uint8_t _pad[3];
} tcr_ew_output;

typedef struct {
  int32_t dtype;        /* compute type: TCR_FLOAT / TCR_DOUBLE / TCR_INT32 / ... */
  int32_t n_inputs, n_outputs, n_instrs;
  int64_t dims[3];      /* D0, D1, D2; n = D0*D1*D2 */
  tcr_ew_input inputs[TCR_EW_MAX_INPUTS];
  tcr_ew_output outputs[TCR_EW_MAX_OUTPUTS];
  tcr_ew_instr instrs[TCR_EW_MAX_INSTRS];
} tcr_ew_program;

int tcr_elementwise(const tcr_ew_program* prog);
/* Map-reduce form: output 0 of the program is not stored but SUMMED over all elements into out[0] (element of the compute type,
 * FLOAT / DOUBLE), then post_op is applied: 0 none, 1 out /= post_imm, 2 out *= post_imm. One launch replaces the reference's
 * elementwise chain + REDUCE_SUM over every rank + scalar DIV of a loss (cfg/tenncor/loss.yml:21-39 -> core.yml:1090-1096:
 * SUB, SQUARE, REDUCE_SUM, DIV by a constant element count). Deterministic: block partials in block order. */
int tcr_elementwise_reduce(const tcr_ew_program* prog, void* out, int post_op, double post_imm);
/* `count` (<= 8) programs of one compute type over the SAME iteration space (equal dims products) in one launch, evaluated
 * in order per element: a program may read, un-broadcast and at the same index, what an earlier one of the call stored.
 * Replaces a chain of elementwise functors whose intermediate results have several readers each (every one of them would be a
 * separate Eigen assignment in the reference: internal/eigen/device.hpp:555-570 calls one per functor).
 * keep[k] = 0 (keep may be NULL = keep all): the result of program k is read only by later programs of this call and is not
 * written to memory (its output pointer still names it). All inputs that come from memory are fetched before the first program
 * runs; results travel between programs through shared memory. */
int tcr_elementwise_multi(const tcr_ew_program* progs, int count, const uint8_t* keep);

/* Backward through one time step of a gated recurrent cell, float32, one launch (a hand-written form of the multi launch above
 * for the shape backprop.hpp:136-142 gives the derivative graph of cfg/tenncor/layer.yml:716-768, every operation rounded on its
 * own like the separate functors):
 *     s   = s_a + s_b                                     gradient reaching the cell's output: two contributions
 *     c   = (c_x * c_y) + (c_z * s)                       gradient of the state: what the next step passes down + this step's
 *     out_k = (x_k * (1 - x_k)) * (y_k * v_k)             SIGMOID gate   (kind 1)
 *     out_k = (1 - x_k * x_k)   * (y_k * v_k)             TANH gate      (kind 2)        v_k = s (sel 0) or c (sel 1)
 * s / c are stored only when their pointer is not NULL. */
typedef struct {
  int64_t n;
  const void *s_a, *s_b;
  const void *c_x, *c_y, *c_z;
  void *s_out, *c_out;
  int32_t n_gates;                 /* <= 6 */
  int32_t kind[6], sel[6];
  const void *x[6], *y[6];
  void* out[6];
} tcr_cell_backward_desc;
int tcr_cell_backward(const tcr_cell_backward_desc* desc);

/* Convenience single-op forms (same kernels, program of length 1). */
int tcr_unary(int opcode, const void* in, void* out, int64_t n, int dtype);
int tcr_binary(int opcode, const void* a, const void* b, void* out, int64_t n, int dtype);
/* n-ary ADD / MUL: out = args[0] op args[1] op ... (operator.hpp:716-731,786-794) */
int tcr_nnary(int opcode, const void* const* args, int nargs, void* out, int64_t n, int dtype);
int tcr_select(const void* cond, const void* then_, const void* else_, void* out, int64_t n, int dtype);
/* CAST (operator.hpp:1239-1260) */
int tcr_cast(const void* in, int in_dtype, void* out, int out_dtype, int64_t n);
/* ASSIGN / ASSIGN_ADD / SUB / MUL / DIV in place on variable storage
 * (eigen::TensAssign, device.hpp:507-544; operator.hpp:1190-1237). `opcode` is the
 * reference opcode TCR_OP_ASSIGN.. */
int tcr_assign(int opcode, void* dst, const void* src, int64_t n, int dtype);
/* RAND_UNIF (operator.hpp:993-1044): out[i] ~ U(lo[i], hi[i]); Philox4x32-10 keyed by
 * (seed, offset + i). Integer dtypes draw from the closed range [lo, hi]. */
int tcr_rand_unif(const void* lo, const void* hi, void* out, int64_t n, int dtype,
                  uint64_t seed, uint64_t offset);
/* The same generator with its (seed, offset) state resident on the device — the stand-in for the
 * reference's process-wide engine (internal/global/random.hpp:78-146). Each call draws n numbers at
 * the current offset and advances it by n on the stream, so a captured CUDA graph produces new
 * numbers at every replay and the sequence equals that of eager calls in the same order. */
int tcr_rand_seed(uint64_t seed, uint64_t offset);
int tcr_rand_unif_stream(const void* lo, const void* hi, void* out, int64_t n, int dtype);

/* ---------------------------------------------------------------- reductions */

/* REDUCE_SUM/PROD/MIN/MAX over the ranks set in `reduce_mask` (bit r = rank r reduced)
 * (operator.hpp:54-132; output keeps rank with 1s, cfg/ops.yml:123-139).
 * `opcode` is TCR_OP_REDUCE_*. */
int tcr_reduce(int opcode, const void* in, void* out, const int64_t shape[TCR_RANK_CAP],
               uint32_t reduce_mask, int dtype);
/* ARGMAX (operator.hpp:136-155): return_dim >= 8 -> flat column-major index of the max
 * of the whole tensor; else index along return_dim. Index stored as the dtype T.
 * Ties resolve to the lowest index. */
int tcr_argmax(const void* in, void* out, const int64_t shape[TCR_RANK_CAP],
               int return_dim, int dtype);

/* ---------------------------------------------------------------- layout ops */

/* Generic coordinate-mapped copy. For every output coordinate c[0..7] the input
 * coordinate is  q[k] = (c[k] * mul[k] + add[k]) / div[k], valid iff the division is
 * exact and 0 <= q[k] < in_shape[k]; invalid -> output element is 0. Covers
 * EXTEND (in dim 1: mul=0), SLICE (add=offset), PAD (add=-pad_lo), STRIDE (mul=incr),
 * SCATTER (div=incr), REVERSE (mul=-1, add=dim-1) (operator.hpp:159-333).
 * `perm[k]` names the input rank that output rank k walks (PERMUTE, operator.hpp:177-210);
 * identity = {0..7}. `elem_size` in bytes (1, 2, 4, 8): layout ops are type-agnostic. */
typedef struct {
  int64_t in_shape[TCR_RANK_CAP];
  int64_t out_shape[TCR_RANK_CAP];
  int32_t perm[TCR_RANK_CAP];
  int64_t mul[TCR_RANK_CAP];
  int64_t add[TCR_RANK_CAP];
  int64_t div[TCR_RANK_CAP];
} tcr_map_desc;

int tcr_map_copy(const void* in, void* out, const tcr_map_desc* desc, int elem_size);

/* Direct forms that build the descriptor (and pick tiled fast paths). */
int tcr_extend(const void* in, void* out, const int64_t in_shape[8], const int64_t bcast[8], int elem_size);
int tcr_permute(const void* in, void* out, const int64_t in_shape[8], const int32_t order[8], int elem_size);
int tcr_slice(const void* in, void* out, const int64_t in_shape[8], const int64_t offsets[8],
              const int64_t extents[8], int elem_size);
int tcr_pad(const void* in, void* out, const int64_t in_shape[8], const int64_t pad_lo[8],
            const int64_t pad_hi[8], int elem_size);
int tcr_stride(const void* in, void* out, const int64_t in_shape[8], const int64_t incrs[8], int elem_size);
int tcr_scatter(const void* in, void* out, const int64_t in_shape[8], const int64_t out_shape[8],
                const int64_t incrs[8], int elem_size);
int tcr_reverse(const void* in, void* out, const int64_t shape[8], uint32_t reverse_mask, int elem_size);
/* CONCAT along `axis` (operator.hpp:336-368): binary concatenation of arbitrary axis
 * extents, or n-ary where every arg has extent 1 along axis. `shapes` = nargs x 8. */
int tcr_concat(const void* const* args, const int64_t* shapes, int nargs, void* out, int axis, int elem_size);
/* `count` strided 2-D copies (rows x row_bytes, byte pitches) in one launch. CONCAT along any rank is one such copy per operand
 * (rows = product of the ranks above the axis, row_bytes = that operand's run along and below it): the per-time-step CONCATs of
 * an unrolled recurrent layer (operator.hpp:336-368, one Eigen assignment each) become one launch. */
typedef struct {
  void* dst;
  const void* src;
  int64_t row_bytes, rows, dst_pitch, src_pitch;
} tcr_copy2d_item;
int tcr_copy2d_batched(const tcr_copy2d_item* items, int count);

/* ------------------------------------------------------------- contractions */

enum { TCR_GEMM_EXACT = 0, /* SIMT FMA in the element type (fp32/fp64/int32) */
       TCR_GEMM_TF32 = 1,  /* tcgen05 kind::tf32, single pass */
       TCR_GEMM_3XTF32 = 2 /* tcgen05 kind::tf32, 3-pass split for fp32-grade accuracy */ };

enum { TCR_EPI_NONE = 0, TCR_EPI_BIAS_N = 1 /* + bias[n] */, TCR_EPI_BIAS_M = 2 /* + bias[m] */ };
/* post-operation on the product, applied last: the chain-rule factor the reference's SIGMOID / TANH gradient rules multiply onto the
 * upstream gradient (tenncor/eteq/backprop.hpp:136-142: MUL(MUL(s, SUB(1, s)), sup), MUL(SUB(1, SQUARE(t)), sup)) when `sup` is this
 * product. `aux` has the layout of C. Only products with k <= 16, unit column stride and n % 4 == 0 accept it (TCR_ERR_UNSUPPORTED else). */
enum { TCR_POST_NONE = 0, TCR_POST_MUL_DSIGMOID = 1 /* c *= aux * (1 - aux) */, TCR_POST_MUL_DTANH = 2 /* c *= 1 - aux^2 */ };

/* C[m,n] (+)= sum_k A(m,k) * B(k,n), fully strided operands:
 *   A(m,k) = a[m*a_sm + k*a_sk + batch*a_sb], likewise B(k,n), C(m,n).
 * MATMUL / CONTRACT 2-D fast path (operator.hpp:1069-1139): in teq shapes
 * C[N,M] = A[K,M] . B[N,K], i.e. row-major (MxK)(KxN). The backward contractions
 * (rank_pairs {{1,1}}, {{0,0}}; tenncor/eteq/backprop.hpp:269-359) are the TN / NT
 * stride variants of the same call, and a trailing PERMUTE{1,0} is absorbed by
 * swapping c_sm / c_sn. */
typedef struct {
  int64_t m, n, k, batch;
  int64_t a_sm, a_sk, a_sb;
  int64_t b_sk, b_sn, b_sb;
  int64_t c_sm, c_sn, c_sb;
  int32_t dtype;
  int32_t precision;   /* TCR_GEMM_* (tensor-core modes: fp32 only) */
  int32_t epilogue;    /* TCR_EPI_* */
  int32_t activation;  /* 0 or TCR_EW_SIGMOID / TCR_EW_TANH applied after bias */
  const void* bias;
  int32_t accumulate;  /* C += (beta = 1) instead of C = */
  int32_t post_op;     /* TCR_POST_* */
  const void* aux;     /* operand of post_op */
} tcr_gemm_desc;

int tcr_gemm(const void* a, const void* b, void* c, const tcr_gemm_desc* desc);

/* Grouped / segmented product for small-batch recurrent steps (gemm_rnn.cu). For g < groups, s < segments:
 *     out_g[m, n] = act_g( sum_s A_s[m, :] . B_{g,s}[:, n] + bias_g[n] )
 * in ONE launch on the tensor cores (FLOAT, TF32 / 3xTF32). It is what the planner lowers these sub-graphs of the
 * reference's unrolled recurrent layers to (cfg/tenncor/layer.yml:716-813):
 *   - the gate products of one time step: CONTRACT(CONCAT(x_t, h_{t-1}), W_g) + EXTEND(b_g) through SIGMOID / TANH for every
 *     gate g — groups = gates, segments = {x_t, h_{t-1}} (the CONCAT, operator.hpp:336-368, is never materialised),
 *     b_trans = 0 (W_g is [K x n], n contiguous);
 *   - the gradient reaching h_{t-1}: ADD over gates of CONTRACT(dpre_g, W_g) (backprop.hpp:269-359 + derive.cpp:49-51) —
 *     groups = 1, segments = gates, b_trans = 1 (row j of W_g is column j of the product);
 *   - optionally the LSTM cell update c_t = cand * in + c_{t-1} * forget, h_t = c_t * out (layer.yml:758-760) in the epilogue.
 * A_s: [m x seg_k[s]] row-major with row pitch a_pitch[s]; B_{g,s}: b_trans == 0 -> [seg_k[s] x n] row-major, b_trans == 1 ->
 * [n x seg_k[s]] row-major, row pitch b_pitch either way; out_g: [m x n] row-major, row pitch out_pitch. Pitches are in
 * elements and multiples of 4, pointers 16-byte aligned (TMA). Split-K runs inside a thread-block cluster and is summed in a
 * fixed order through distributed shared memory: deterministic, no workspace. */
typedef struct {
  int64_t m, n;
  int32_t groups;   /* 1, 2 or 4 */
  int32_t segments; /* 1..4; groups * segments <= 8 */
  int64_t seg_k[4];
  const void* a[4];
  int64_t a_pitch[4];
  const void* b[4][4]; /* [group][segment] */
  int64_t b_pitch;
  int32_t b_trans;
  int32_t precision;   /* TCR_GEMM_TF32 | TCR_GEMM_3XTF32 */
  const void* bias[4]; /* per group: n elements, or null */
  int32_t act[4];      /* per group: 0 | TCR_EW_SIGMOID | TCR_EW_TANH */
  void* out[4];        /* per group; may be null only when `cell` is set (activation not needed afterwards) */
  int64_t out_pitch;
  int32_t accumulate;  /* out += instead of out = */
  int32_t cell;        /* 1: LSTM cell epilogue (groups == 4) */
  int32_t role_cand, role_in, role_forget, role_out; /* which group is which gate */
  const void* c_prev;  /* [m x n] pitch state_pitch, or null for a zero state */
  void* c_out;
  void* h_out;
  int64_t state_pitch;
} tcr_gemm_group_desc;

int tcr_gemm_grouped(const tcr_gemm_group_desc* desc);
/* `count` CONSECUTIVE TIME STEPS of a recurrent layer in ONE launch: descs[t] is the gate launch of step t (cell = 1), step
 * t + 1 reads h_out of step t as one of its A segments and c_out of step t as its c_prev; weights, biases and extents are
 * those of descs[0]. The CTAs stay resident over the whole sequence (tensor maps, barriers and TMEM set up once, weight tiles
 * of step t + 1 prefetched during the epilogue of step t) and meet at a grid barrier between steps — the reference runs
 * 4 x seq Eigen contractions + their elementwise tails one after the other (cfg/tenncor/layer.yml:716-768). prepare() builds
 * the per-step tables in device memory once (not capturable); launch() is one kernel launch (capturable); all clusters must fit
 * the device at once (prepare fails otherwise, and the caller keeps the per-step launches). */
int tcr_gemm_grouped_seq_prepare(const tcr_gemm_group_desc* descs, int count, void** handle);
int tcr_gemm_grouped_seq_launch(void* handle);
int tcr_gemm_grouped_seq_destroy(void* handle);
/* host-only validation of everything but the pointers (the planner asks before it commits to this lowering) */
int tcr_gemm_grouped_check(const tcr_gemm_group_desc* desc);
/* profiling aid (TCR_RNN_DEBUG=1): SM-clock stamps of the first CTA of the last tcr_gemm_grouped launch */
int tcr_rnn_debug_read(long long out[16]);

/* General CONTRACT (operator.hpp:1069-1101, shape rule cfg/ops.yml:547-565): pairs
 * (a_rank, b_rank) are contracted; out dims = b-free (in order) then a-free. Used when
 * the operands cannot be viewed as strided matrices. */
int tcr_contract(const void* a, const void* b, void* out, const int64_t a_shape[8],
                 const int64_t b_shape[8], const int32_t* pairs, int npairs, int dtype);

/* CONV (operator.hpp:1143-1187): N-d *valid correlation* (no flip, no stride,
 * golden internal/eigen/test/test_operator.cpp:2630-2634). Kernel rank i slides along
 * image rank order[i]; ranks not listed are appended in order. */
int tcr_conv(const void* image, const void* kernel, void* out, const int64_t img_shape[8],
             const int64_t kern_shape[8], const int32_t order[8], int dtype);

/* Patch gather for the conv2d composite (cfg/tenncor/nn.yml:48-98 = PAD along a fresh rank +
 * REVERSE(kernel) + CONV + PERMUTE): cols[pos][win] = image[pos + win], where `pos` runs over the
 * valid positions (img_shape - win_shape + 1 per rank, rank 0 fastest) and `win` over the window
 * coordinates (rank 0 fastest); each row is zero-filled up to `row_pitch` elements (a multiple
 * of 4). The multi-filter correlation is then ONE tcr_gemm over `cols` on the tensor cores
 * instead of operator.hpp:1143-1187's scalar slide over a mostly-zero rank. 4-byte elements. */
int tcr_im2col(const void* image, void* cols, const int64_t img_shape[8], const int64_t win_shape[8],
               int64_t row_pitch, int elem_size);
/* tcr_im2col + tcr_gemm(cols, b, c) WITHOUT the patch matrix in memory (implicit GEMM): the tensor-core kernel's producer warp gathers
 * each 128-row x 32-k tile of `cols` straight from the image with cp.async into the shared-memory arrangement the MMA reads, so the
 * matrix (9x the image for a 3 x 3 window) is neither written nor re-read. `desc` describes the product as if `a` were the patch
 * matrix with row pitch k (a_sm / a_sk are ignored); bias / activation epilogues as in tcr_gemm. Eligible views: image [C, W, H,
 * images...], window [C, kw, kh, 1...] with (kw * C) % 32 == 0, FLOAT, TF32 / 3xTF32. Anything else returns TCR_ERR_UNSUPPORTED
 * without launching (callers keep the two-call form). Replaces the reference's scalar slide over a padded rank for conv2d
 * (internal/eigen/operator.hpp:1143-1187 under cfg/tenncor/nn.yml:48-98). */
int tcr_gemm_patches(const void* image, const void* b, void* c, const tcr_gemm_desc* desc, const int64_t img_shape[8],
                     const int64_t win_shape[8]);
/* Adjoint of tcr_im2col (the conv2d image gradient: cols = upstream . kernel^T on the tensor cores, then this):
 * image[u] = sum over window coordinates w with a valid position u - w of cols[position(u - w)][w]. Every image
 * element is written (zero where no window reaches). Terms are added in a fixed order: deterministic. FLOAT only. */
int tcr_col2im(const void* cols, void* image, const int64_t img_shape[8], const int64_t win_shape[8],
               int64_t row_pitch, int dtype);

/* --------------------------------------------------------------- collectives */

/* NCCL data-parallel gradient exchange (replaces tenncor/distr's gRPC/consul path for
 * batch-sharded training, tenncor/eteq/opsvc/service.hpp:110-160). One process per
 * GPU; rank 0 creates the id, the launcher ships it to the other ranks. */
#define TCR_COMM_ID_BYTES 128
int tcr_comm_unique_id(char id[TCR_COMM_ID_BYTES]);
int tcr_comm_init(int rank, int nranks, const char id[TCR_COMM_ID_BYTES]);
int tcr_comm_destroy(void);
int tcr_comm_rank(void);
int tcr_comm_size(void); /* 1 when no communicator */
/* in-place SUM all-reduce followed by `scale` (1/nranks for mean-type losses). FLOAT buffers inside the symmetric region
 * (tcr_comm_symm_alloc) are exchanged by ONE kernel over NVLink peer memory (allreduce_p2p.cu: every rank maps every peer's
 * region through CUDA IPC; one-shot below 512 KB; above, every rank pulls and sums its slice and pushes it to all peers; rank-order sums, bit-identical on all
 * ranks); everything else goes through ncclAllReduce. */
int tcr_allreduce_sum(void* buf, int64_t n, int dtype, double scale);
/* Symmetric region: memory every peer can address. Ranks must allocate the same sizes in the same order (the planner does: all
 * ranks build the same plan). TCR_ERR_UNSUPPORTED when no peer path exists or the region is full: use tcr_alloc + NCCL then. */
int tcr_comm_symm_alloc(void** out, size_t bytes);
int tcr_comm_symm_reset(void); /* forget every symmetric allocation (call on all ranks together, with no plan alive) */
int tcr_comm_p2p_ready(void);  /* 1 when the peer-memory path is active */

#ifdef __cplusplus
}
#endif
#endif /* TCR_B200_H */
