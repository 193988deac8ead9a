"""-m "not gpu": the C-ABI library loads and exports every symbol include/graphchem_b200.h declares."""
import ctypes
import os
import re

from graphchem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "graphchem_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 12
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/graphchem_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding table and header disagree"


def test_abi_version_and_layout():
    lib = _lib.lib()
    assert lib.gcb_abi_version() == 3
    assert b"sm_100a" in lib.gcb_build_info()
    cfg = _lib.ModelCfg(42, 29, 1, 128, 3, 3, 128, 0, 0, 0, 0.0, 0)
    offs = (ctypes.c_int64 * 65)()
    n = lib.gcb_param_layout(ctypes.byref(cfg), offs)
    assert n == 16
    assert offs[n] == 124673  # parameter count printed in examples/predict_cn.ipynb:181-196
    cfg2 = _lib.ModelCfg(64, 32, 1, 128, 4, 2, 64, 0, 0, 1, 0.0, 0)
    n2 = lib.gcb_param_layout(ctypes.byref(cfg2), offs)
    assert n2 == 14 and offs[n2] == 90689
    assert lib.gcb_workspace_bytes(ctypes.byref(cfg2), 23 * 16, 50 * 16, 16, 3) > 0
    bad = _lib.ModelCfg(64, 32, 1, 100, 4, 2, 64, 0, 0, 1, 0.0, 0)  # D not a multiple of 8
    assert lib.gcb_workspace_bytes(ctypes.byref(bad), 10, 20, 1, 0) == _lib.GCB_ERR_UNSUPPORTED


def test_argument_validation_without_a_gpu():
    """Entry points reject malformed arguments before touching CUDA (status codes, never exceptions): the new config
    values of round 2 (aggr = max, pool = mean) are accepted by the sizing helpers, unknown ones are refused, and
    gcb_train_steps refuses missing buffers."""
    lib = _lib.lib()
    ok = _lib.ModelCfg(64, 32, 1, 128, 4, 2, 64, 0, _lib.AGGR_MAX, 1, 0.0, _lib.POOL_MEAN)
    assert lib.gcb_workspace_bytes(ctypes.byref(ok), 230, 500, 10, 3) > 0
    add = _lib.ModelCfg(64, 32, 1, 128, 4, 2, 64, 0, _lib.AGGR_ADD, 1, 0.0, _lib.POOL_ADD)
    # max aggregation keeps per-edge messages instead of the aggregated [N, 2D] operands: a different workspace
    assert lib.gcb_workspace_bytes(ctypes.byref(ok), 230, 500, 10, 3) != lib.gcb_workspace_bytes(ctypes.byref(add), 230, 500, 10, 3)
    for bad in (_lib.ModelCfg(64, 32, 1, 128, 4, 2, 64, 0, 7, 1, 0.0, 0), _lib.ModelCfg(64, 32, 1, 128, 4, 2, 64, 0, 0, 1, 0.0, 9)):
        assert lib.gcb_workspace_bytes(ctypes.byref(bad), 230, 500, 10, 3) < 0
    rc = lib.gcb_train_steps(ctypes.byref(ok), 0, None, None, None, None, None, None, 0, 0, None, None, None, None, None, None,
                             None, 0.9, 0.999, 1e-8, None, None)
    assert rc == _lib.GCB_ERR_INVALID_ARG
