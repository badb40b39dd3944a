"""The C-ABI library loads without a GPU, exports every symbol the header declares, and
refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tcr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tcr_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from tenncor_b200 import cabi
    assert header_symbols() == sorted(cabi.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    from tenncor_b200 import cabi
    lib = cabi.lib()
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_struct_layouts_match_header(built):
    from tenncor_b200 import cabi
    # sizes implied by include/tcr_b200.h on LP64
    assert C.sizeof(cabi.EwInstr) == 16
    assert C.sizeof(cabi.EwInput) == 16
    assert C.sizeof(cabi.EwOutput) == 16
    assert C.sizeof(cabi.EwProgram) == 16 + 24 + 8 * 16 + 4 * 16 + 32 * 16
    assert C.sizeof(cabi.MapDesc) == 8 * 8 * 5 + 8 * 4
    assert C.sizeof(cabi.GemmDesc) == 13 * 8 + 4 * 4 + 8 + 8 + 8  # + post_op (in the old pad) + aux pointer


def test_no_cpu_fallback_without_device(built):
    from tenncor_b200 import cabi
    lib = cabi.lib()
    if lib.tcr_device_count() > 0:
        pytest.skip("a GPU is present")
    assert lib.tcr_init(0) != 0
    assert b"no CPU fallback" in lib.tcr_last_error()
    p = C.c_void_p()
    assert lib.tcr_alloc(C.byref(p), C.c_size_t(16)) != 0
    with pytest.raises(cabi.TcrError):
        cabi.init(0)
    buf = (C.c_float * 4)()
    assert lib.tcr_unary(cabi.OP["EXP"], buf, buf, C.c_int64(4), cabi.FLOAT) != 0
