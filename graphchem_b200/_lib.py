"""ctypes binding of the C-ABI library (include/graphchem_b200.h).

The library is built in-tree (graphchem_b200/lib/libgraphchem_b200.so) by `build()` below or by
`__graft_entry__.build()`.  There is NO fallback: if the library is missing the product raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "lib", "libgraphchem_b200.so")

GCB_OK = 0
GCB_ERR_INVALID_ARG = -1
GCB_ERR_INDEX = -2
GCB_ERR_UNSUPPORTED = -3
GCB_ERR_CUDA = -4
GCB_ERR_WORKSPACE = -5

ACT_SOFTPLUS, ACT_SIGMOID, ACT_RELU, ACT_TANH, ACT_LEAKY_RELU = range(5)
AGGR_ADD, AGGR_MEAN, AGGR_MAX = 0, 1, 2
POOL_ADD, POOL_MEAN = 0, 1
F32, BF16 = 0, 1
TRAINING, SAVE_FOR_BWD = 1, 2

PACK_HEADER_INTS = 48
(H_MAGIC, H_N, H_E, H_G, H_M, H_AV, H_BV, H_TOTAL_INTS, H_X_TOK, H_E_TOK, H_SRC, H_DST, H_IN_PTR, H_IN_EID,
 H_IN_SRC, H_OUT_PTR, H_OUT_EID, H_OUT_DST, H_GRAPH_PTR, H_GRAPH_NODE, H_ATOK_PTR, H_ATOK_NODE, H_BTOK_PTR,
 H_BTOK_ROW, H_NODE_GRAPH, H_IN_SRC4, H_IN_EID4, H_OUT_DST4, H_OUT_INPOS4, H_OUT_INPOS, H_TILE_PTR,
 H_NTILES, H_FLAGS, H_BOND_POS, H_COMPACT_INTS) = range(35)
PACK_EACH = 1
TILE_ROWS = 128
PACK_MAGIC = 0x47434232
RAW_HEADER_INTS = 16
MAX_PARAM_TENSORS = 64


class ModelCfg(C.Structure):
    """Mirror of `gcb_model_cfg` (constructor arguments of MoleculeGCN, reference gcn.py:35-42)."""
    _fields_ = [("atom_vocab", C.c_int32), ("bond_vocab", C.c_int32), ("out_dim", C.c_int32), ("D", C.c_int32),
                ("L", C.c_int32), ("n_readout", C.c_int32), ("R", C.c_int32), ("act", C.c_int32),
                ("aggr", C.c_int32), ("dtype", C.c_int32), ("p_dropout", C.c_float), ("pool", C.c_int32)]


class GcbError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into graphchem_b200/lib (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise GcbError(f"building {LIB_PATH} failed (exit {res.returncode})")
    return LIB_PATH


_lib: Optional[C.CDLL] = None
_lock = threading.Lock()

_vp, _i32, _i64, _u64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
_cfgp = C.POINTER(ModelCfg)

# name -> (restype, argtypes); one entry per declaration in include/graphchem_b200.h
SIGNATURES = {
    "gcb_pack_ints": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "gcb_pack_ints_flags": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "gcb_collate_pack_each_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "gcb_pack_host": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64]),
    "gcb_collate_pack_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "gcb_raw_ints": (_i64, [_i32, _i32, _i32]),
    "gcb_collate_raw_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp]),
    "gcb_pack_device_scratch_ints": (_i64, [_vp]),
    "gcb_pack_device": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "gcb_pack_expand_ell": (C.c_int, [_vp, _vp, _vp]),
    "gcb_debug_pack_emulate_host": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _i32]),
    "gcb_param_layout": (C.c_int, [_cfgp, _vp]),
    "gcb_derived_bytes": (_i64, [_cfgp]),
    "gcb_prepare_weights": (C.c_int, [_cfgp, _vp, _vp, _vp]),
    "gcb_workspace_bytes": (_i64, [_cfgp, _i32, _i32, _i32, _i32]),
    "gcb_forward": (C.c_int, [_cfgp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _u64, _vp, _vp, _vp, _vp]),
    "gcb_backward": (C.c_int, [_cfgp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _u64, _vp, _vp, _vp, _vp, _i32, _vp]),
    "gcb_mse_loss": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _f32, _vp]),
    "gcb_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _f32, _f32, _f32, _vp]),
    "gcb_train_steps": (C.c_int, [_cfgp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _f32, _f32, _f32, _vp, _vp]),
    "gcb_debug_gemm_nt_bf16": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _i32, _vp]),
    "gcb_debug_gemm_tn_bf16": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "gcb_launch_count": (_i64, []),
    "gcb_timing_enable": (C.c_int, [_i32]),
    "gcb_timing_report": (_i64, [_vp, _i64]),
    "gcb_abi_version": (C.c_int, []),
    "gcb_last_cuda_error": (C.c_char_p, []),
    "gcb_build_info": (C.c_char_p, []),
}


def lib() -> C.CDLL:
    """Load the C-ABI library.  Raises GcbError (never falls back) when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise GcbError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or `make -C graphchem_b200/csrc`).  graphchem_b200 has no CPU or PyTorch fallback.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
            _lib = handle
    return _lib


def check(rc: int, what: str = "") -> int:
    """Translate a C-ABI status into the exception the reference would raise."""
    if rc >= 0:
        return rc
    if rc == GCB_ERR_INDEX:
        raise IndexError(f"{what}: index out of range (edge_index must address rows of the bond matrix, i.e. "
                         "max(edge_index) < min(num_nodes, num_edges) when n_messages >= 1; tokens must be "
                         "< vocab size)")
    if rc == GCB_ERR_INVALID_ARG:
        raise ValueError(f"{what}: invalid argument")
    if rc == GCB_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: configuration not supported by the sm_100a kernels "
                                  "(embedding_dim must be a multiple of 8; aggr in {'add','mean','max'})")
    if rc == GCB_ERR_CUDA:
        raise GcbError(f"{what}: CUDA error: {lib().gcb_last_cuda_error().decode()}")
    if rc == GCB_ERR_WORKSPACE:
        raise GcbError(f"{what}: workspace too small")
    raise GcbError(f"{what}: error {rc}")
