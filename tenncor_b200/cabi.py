"""ctypes binding of the C-ABI in include/tcr_b200.h (libtcr_b200.so).

This is the same boundary the C++ host (`tenncor_b200/host`) links against; the Python
binding exists so that tests and bench.py can drive individual kernels with plain device
pointers. There is no CPU fallback: every call raises `TcrError` when the library or a
CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtcr_b200.so")

RANK_CAP = 8
EW_MAX_INPUTS, EW_MAX_OUTPUTS, EW_MAX_INSTRS, EW_NREGS = 8, 4, 32, 8
COMM_ID_BYTES = 128

# egen::_GENERATED_DTYPE (cfg/fulltype.yml)
DOUBLE, FLOAT, INT8, UINT8, INT16, UINT16, INT32, UINT32, INT64, UINT64 = range(1, 11)
NP_DTYPE = {DOUBLE: np.float64, FLOAT: np.float32, INT8: np.int8, UINT8: np.uint8, INT16: np.int16,
            UINT16: np.uint16, INT32: np.int32, UINT32: np.uint32, INT64: np.int64, UINT64: np.uint64}
DTYPE_OF = {np.dtype(v): k for k, v in NP_DTYPE.items()}

# egen::_GENERATED_OPCODE (cfg/ops.yml order)
OPCODES = [
    "BAD_OP", "IDENTITY", "ABS", "NEG", "SIN", "COS", "TAN", "EXP", "LOG", "SQRT", "ROUND",
    "SIGMOID", "TANH", "SQUARE", "CUBE", "RAND_UNIF", "REVERSE", "REDUCE_SUM", "REDUCE_PROD",
    "REDUCE_MIN", "REDUCE_MAX", "ARGMAX", "PERMUTE", "EXTEND", "RESHAPE", "SLICE", "PAD",
    "STRIDE", "SCATTER", "POW", "ADD", "SUB", "MUL", "DIV", "MIN", "MAX", "EQ", "NEQ", "LT",
    "GT", "MATMUL", "CONTRACT", "CONV", "SELECT", "CONCAT", "ASSIGN", "ASSIGN_ADD",
    "ASSIGN_SUB", "ASSIGN_MUL", "ASSIGN_DIV", "CAST",
]
OP = {n: i for i, n in enumerate(OPCODES)}
EW_MOV, EW_CONST = 64, 65
GEMM_EXACT, GEMM_TF32, GEMM_3XTF32 = 0, 1, 2
EPI_NONE, EPI_BIAS_N, EPI_BIAS_M = 0, 1, 2
POST_NONE, POST_MUL_DSIGMOID, POST_MUL_DTANH = 0, 1, 2

# every symbol include/tcr_b200.h declares (checked by tests/test_cabi_symbols.py)
SYMBOLS = [
    "tcr_init", "tcr_shutdown", "tcr_last_error", "tcr_device_count", "tcr_sm_count", "tcr_stream",
    "tcr_sync", "tcr_alloc", "tcr_free", "tcr_arena_stats", "tcr_arena_trim", "tcr_host_alloc",
    "tcr_host_free", "tcr_h2d", "tcr_h2d_prefetch", "tcr_prefetch_commit", "tcr_prefetch_sync", "tcr_d2h", "tcr_d2d", "tcr_memset", "tcr_event_create",
    "tcr_event_destroy", "tcr_event_record", "tcr_event_sync", "tcr_event_elapsed_ms", "tcr_graph_begin", "tcr_graph_lane", "tcr_graph_record", "tcr_graph_wait",
    "tcr_graph_end", "tcr_graph_launch", "tcr_graph_destroy", "tcr_launch_count", "tcr_elementwise", "tcr_elementwise_reduce", "tcr_elementwise_multi", "tcr_cell_backward",
    "tcr_unary", "tcr_binary", "tcr_nnary", "tcr_select", "tcr_cast", "tcr_assign", "tcr_rand_unif", "tcr_rand_seed", "tcr_rand_unif_stream",
    "tcr_reduce", "tcr_argmax", "tcr_map_copy", "tcr_extend", "tcr_permute", "tcr_slice", "tcr_pad",
    "tcr_stride", "tcr_scatter", "tcr_reverse", "tcr_concat", "tcr_copy2d_batched", "tcr_gemm", "tcr_gemm_grouped", "tcr_gemm_grouped_check", "tcr_gemm_grouped_seq_prepare", "tcr_gemm_grouped_seq_launch", "tcr_gemm_grouped_seq_destroy", "tcr_rnn_debug_read", "tcr_contract", "tcr_conv", "tcr_im2col", "tcr_gemm_patches", "tcr_col2im",
    "tcr_comm_unique_id", "tcr_comm_init", "tcr_comm_destroy", "tcr_comm_rank", "tcr_comm_size",
    "tcr_allreduce_sum", "tcr_comm_symm_alloc", "tcr_comm_symm_reset", "tcr_comm_p2p_ready",
]


class TcrError(RuntimeError):
    pass


class EwInstr(C.Structure):
    _fields_ = [("op", C.c_uint8), ("dst", C.c_uint8), ("a", C.c_uint8), ("b", C.c_uint8),
                ("c", C.c_uint8), ("_pad", C.c_uint8 * 3), ("imm", C.c_double)]


class EwInput(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("bcast", C.c_uint8 * 3), ("_pad", C.c_uint8)]


class EwOutput(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("reg", C.c_uint8), ("_pad", C.c_uint8 * 3)]


class EwProgram(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("n_inputs", C.c_int32), ("n_outputs", C.c_int32),
                ("n_instrs", C.c_int32), ("dims", C.c_int64 * 3),
                ("inputs", EwInput * EW_MAX_INPUTS), ("outputs", EwOutput * EW_MAX_OUTPUTS),
                ("instrs", EwInstr * EW_MAX_INSTRS)]


class MapDesc(C.Structure):
    _fields_ = [("in_shape", C.c_int64 * 8), ("out_shape", C.c_int64 * 8), ("perm", C.c_int32 * 8),
                ("mul", C.c_int64 * 8), ("add", C.c_int64 * 8), ("div", C.c_int64 * 8)]


class GemmDesc(C.Structure):
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64), ("batch", C.c_int64),
                ("a_sm", C.c_int64), ("a_sk", C.c_int64), ("a_sb", C.c_int64),
                ("b_sk", C.c_int64), ("b_sn", C.c_int64), ("b_sb", C.c_int64),
                ("c_sm", C.c_int64), ("c_sn", C.c_int64), ("c_sb", C.c_int64),
                ("dtype", C.c_int32), ("precision", C.c_int32), ("epilogue", C.c_int32),
                ("activation", C.c_int32), ("bias", C.c_void_p), ("accumulate", C.c_int32),
                ("post_op", C.c_int32), ("aux", C.c_void_p)]


class CellBackwardDesc(C.Structure):
    """tcr_cell_backward_desc (include/tcr_b200.h)"""
    _fields_ = [("n", C.c_int64), ("s_a", C.c_void_p), ("s_b", C.c_void_p), ("c_x", C.c_void_p), ("c_y", C.c_void_p), ("c_z", C.c_void_p),
                ("s_out", C.c_void_p), ("c_out", C.c_void_p), ("n_gates", C.c_int32), ("kind", C.c_int32 * 6), ("sel", C.c_int32 * 6),
                ("x", C.c_void_p * 6), ("y", C.c_void_p * 6), ("out", C.c_void_p * 6)]


class GemmGroupDesc(C.Structure):
    """tcr_gemm_group_desc (include/tcr_b200.h)"""
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("groups", C.c_int32), ("segments", C.c_int32),
                ("seg_k", C.c_int64 * 4), ("a", C.c_void_p * 4), ("a_pitch", C.c_int64 * 4),
                ("b", (C.c_void_p * 4) * 4), ("b_pitch", C.c_int64), ("b_trans", C.c_int32), ("precision", C.c_int32),
                ("bias", C.c_void_p * 4), ("act", C.c_int32 * 4), ("out", C.c_void_p * 4), ("out_pitch", C.c_int64),
                ("accumulate", C.c_int32), ("cell", C.c_int32),
                ("role_cand", C.c_int32), ("role_in", C.c_int32), ("role_forget", C.c_int32), ("role_out", C.c_int32),
                ("c_prev", C.c_void_p), ("c_out", C.c_void_p), ("h_out", C.c_void_p), ("state_pitch", C.c_int64)]


_lib = None


def lib():
    """Load libtcr_b200.so (no device needed to load; compute calls need tcr_init)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TcrError("libtcr_b200.so is not built (%s missing): run "
                           "`python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.tcr_last_error.restype = C.c_char_p
        _lib.tcr_stream.restype = C.c_void_p
        _lib.tcr_launch_count.restype = C.c_uint64
    return _lib


def check(rc):
    if rc != 0:
        raise TcrError("tcr_b200 error %d: %s" % (rc, lib().tcr_last_error().decode()))


_inited = False


def init(device=None):
    global _inited
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    check(lib().tcr_init(int(device)))
    _inited = True


def shape8(shape):
    s = [int(d) for d in shape][:8]
    return (C.c_int64 * 8)(*(s + [1] * (8 - len(s))))


class DeviceBuffer:
    """A device allocation from the library arena, freed on garbage collection."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib().tcr_alloc(C.byref(p), C.c_size_t(max(self.nbytes, 1))))
        self.ptr = p.value

    def __del__(self):
        try:
            if getattr(self, "ptr", None) and _lib is not None:
                _lib.tcr_free(C.c_void_p(self.ptr))
        except Exception:
            pass
        self.ptr = None


def to_device(arr):
    arr = np.ascontiguousarray(arr)
    buf = DeviceBuffer(arr.nbytes)
    check(lib().tcr_h2d(C.c_void_p(buf.ptr), arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes)))
    check(lib().tcr_sync())  # the pageable source may be released by the caller
    return buf


def empty(n, dtype):
    return DeviceBuffer(int(n) * np.dtype(dtype).itemsize)


def to_host(buf, n, dtype):
    out = np.empty(int(n), dtype=dtype)
    check(lib().tcr_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(buf.ptr), C.c_size_t(out.nbytes)))
    check(lib().tcr_sync())
    return out


def sync():
    check(lib().tcr_sync())


def make_program(dtype, dims, inputs, outputs, instrs):
    """inputs: [(ptr, dtype, (b0, b1, b2))]; outputs: [(ptr, dtype, reg)];
    instrs: [(op, dst, a, b, c, imm)]"""
    p = EwProgram()
    p.dtype = dtype
    p.n_inputs, p.n_outputs, p.n_instrs = len(inputs), len(outputs), len(instrs)
    d = list(dims) + [1] * (3 - len(dims))
    for k in range(3):
        p.dims[k] = int(d[k])
    for k, (ptr, dt, bc) in enumerate(inputs):
        p.inputs[k].ptr = ptr
        p.inputs[k].dtype = dt
        for j in range(3):
            p.inputs[k].bcast[j] = int(bc[j])
    for k, (ptr, dt, reg) in enumerate(outputs):
        p.outputs[k].ptr = ptr
        p.outputs[k].dtype = dt
        p.outputs[k].reg = reg
    for k, ins in enumerate(instrs):
        op, dst, a, b, c, imm = (list(ins) + [0, 0, 0, 0.0])[:6]
        p.instrs[k].op, p.instrs[k].dst, p.instrs[k].a = op, dst, a
        p.instrs[k].b, p.instrs[k].c, p.instrs[k].imm = b, c, float(imm)
    return p
