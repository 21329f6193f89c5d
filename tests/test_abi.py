"""The C-ABI library without a GPU: it loads, exports every symbol include/mindrl_b200.h declares,
the ctypes table mirrors the header, argument errors come back as codes, and compute entry points
fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mindrl_b200.h")


@pytest.fixture(scope="module")
def lib():
    from mindrl_b200 import build
    build.build()
    from mindrl_b200 import _ffi
    return _ffi.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from mindrl_b200 import _ffi
    names = declared_symbols()
    assert len(names) >= 45
    out = subprocess.check_output(["nm", "-D", "--defined-only", _ffi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (mrl_[a-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(_ffi.SIGNATURES), (set(names) ^ set(_ffi.SIGNATURES))
    for n in names:
        assert getattr(lib, n) is not None


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "mindrl_b200.h"\n#include "mrl_arith.h"\nint main(void){return MRL_OK + (int)mrl_detpowf(1.0f, 2.0f) - 1;}\n')
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(tmp_path / "t"), "-lm"])
    subprocess.check_call([str(tmp_path / "t")])


def test_library_has_blackwell_code_only():
    from mindrl_b200 import _ffi
    out = subprocess.run(["cuobjdump", "-lelf", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out)


def test_argument_errors_are_codes(lib):
    from mindrl_b200 import _ffi
    assert lib.mrl_version() >= 100
    h = C.c_int64(0)
    rb = np.asarray([4], dtype=np.int64)
    assert lib.mrl_urb_create(0, 1, rb.ctypes.data, None, 0, C.byref(h)) == _ffi.MRL_ERR_INVALID
    assert h.value == -1 and b"capacity" in lib.mrl_last_error()
    assert lib.mrl_prb_create(10, 0.5, 0, rb.ctypes.data, 0, 0, C.byref(h)) == _ffi.MRL_ERR_INVALID
    assert lib.mrl_urb_destroy(12345) == _ffi.MRL_ERR_HANDLE
    assert lib.mrl_prb_destroy(12345) == _ffi.MRL_ERR_HANDLE
    assert lib.mrl_prb_update(777, None, None, 4, None) == _ffi.MRL_ERR_HANDLE
    assert lib.mrl_discounted_return(None, None, None, None, 4, 4, 1, 1.5, 0, 0, None) == _ffi.MRL_ERR_INVALID
    assert b"range [0, 1]" in lib.mrl_last_error()
    assert lib.mrl_discounted_return(None, None, None, None, 4, 4, 2, 0.9, 1, 0, None) == _ffi.MRL_ERR_INVALID
    assert lib.mrl_set_option(b"no_such_option", 1) == _ffi.MRL_ERR_INVALID
    assert lib.mrl_set_option(b"gather_path", 0) == _ffi.MRL_OK
    assert lib.mrl_launch_count() >= 0


def test_no_cpu_fallback(lib):
    """On a box without a CUDA device the product path must fail loudly, not compute on the CPU."""
    from mindrl_b200 import _ffi
    n = C.c_int(0)
    rc = lib.mrl_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    import mindrl_b200 as M
    with pytest.raises(_ffi.MindRLError) as e:
        M.PriorityReplayBuffer(0.6, 64, 4, [(4,)], [np.float32])
    assert e.value.code == _ffi.MRL_ERR_CUDA
    with pytest.raises(_ffi.MindRLError):
        M.DiscountedReturn(0.99)(np.ones((4, 2), np.float32), np.zeros((4, 2), bool), np.zeros(2, np.float32))


def test_product_never_imports_the_oracle():
    """nothing under mindrl_b200/ may import, include, link or dlopen anything under oracle/"""
    pkg = os.path.join(ROOT, "mindrl_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.+oracle|oracle/|libmrl_oracle|\borc_[a-z]")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), os.path.join(dirpath, f)
    from mindrl_b200 import _ffi
    out = subprocess.run(["ldd", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


AOT_HEADER = os.path.join(ROOT, "include", "mindrl_b200_aot.h")


def test_aot_header_symbols_are_exported(lib):
    """every operator of include/mindrl_b200_aot.h exists as <Func> and <Func>Init -- the two symbols MindSpore's
    ops.Custom("<so>:<Func>", ..., "aot") binds (reference: utils/mcts/cpu/cpu_mcts_ops.cc:38-53) -- and so do the
    mrl_aot_extra_* helpers; the ctypes table mirrors the header"""
    from mindrl_b200 import _ffi
    text = re.sub(r"/\*.*?\*/", "", open(AOT_HEADER).read(), flags=re.S)
    ops = re.findall(r"^MRL_AOT_OP\((\w+)\);", text, flags=re.M)
    helpers = sorted(set(re.findall(r"\b(mrl_aot_[a-z0-9_]+)\s*\(", text)))
    assert sorted(ops) == sorted(_ffi.AOT_OPS) and len(ops) == 10
    out = subprocess.check_output(["nm", "-D", "--defined-only", _ffi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (\w+)", out))
    for op in ops:
        assert op in exported and op + "Init" in exported, op
    assert set(helpers) <= exported
    assert set(helpers) | set(ops) | {o + "Init" for o in ops} == set(_ffi.AOT_SIGNATURES)


def test_aot_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "mindrl_b200_aot.h"\nint main(void){return 0;}\n')
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


def test_aot_attribute_reader_and_init_without_a_gpu(lib):
    """the Init half of an AOT operator only reads attributes and shapes: it runs without a device.  The stand-in
    AotExtra hands back what was registered; a bad registration is the reference's error code 2; the operator itself
    fails loudly (2) when there is no device instead of computing anywhere else"""
    from aot_host import AotOp
    n = C.c_int(0)
    no_gpu = not (lib.mrl_device_count(C.byref(n)) == 0 and n.value > 0)
    good = AotOp("PriorityReplayBufferCreate", capacity=100, alpha=0.6, shapes=[[4], [1]], dtypes="float32,int32",
                 seed0=1, seed1=2)
    h = np.full(1, -5, dtype=np.int64)
    rc = good([(h.ctypes.data, (1,), np.int64)])
    assert good.init_rc == 0
    if no_gpu:
        assert rc == 2 and h[0] == -5
    bad = AotOp("PriorityReplayBufferCreate", capacity=100, alpha=0.6, shapes=[[4], [1]], dtypes="float32",
                seed0=1, seed1=2)
    assert bad([(h.ctypes.data, (1,), np.int64)]) == 2 and bad.init_rc == 2
    assert b"shapes and dtypes" in lib.mrl_last_error()
    weird = AotOp("GaeScan", gamma=0.99, **{"lambda": 0.95}, layout=7)
    x = np.zeros((4, 4), np.float32)
    p = (x.ctypes.data, x.shape, np.float32)
    d = (x.ctypes.data, x.shape, np.bool_)
    assert weird([p, p, p, d, p, p]) == 2
