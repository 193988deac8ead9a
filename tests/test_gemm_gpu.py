"""-m gpu: the tcgen05/TMEM/TMA GEMM (gemm_tc.cuh) against a torch fp32 reference of the same op
(inputs rounded to bf16, fp32 accumulation, output rounded to bf16): tolerance 2^-7 relative to
the row scale, i.e. one bf16 ulp of the result plus accumulation-order noise."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from graphchem_b200 import _lib

pytestmark = pytest.mark.gpu

ACT = {None: -1, "softplus": 0, "relu": 2, "tanh": 3}


def _run(rows, N, K, bias=False, deg=False, res=False, act=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = (torch.randn(rows, K, generator=g) * 0.5).bfloat16().cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().cuda()
    Cm = torch.full((rows, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    b = torch.randn(N, generator=g).cuda() if bias else None
    rp = None
    if deg:
        d = torch.randint(0, 4, (rows,), generator=g)
        rp = torch.cat([torch.zeros(1, dtype=torch.long), d.cumsum(0)]).int().cuda()
    R = torch.randn(rows, N, generator=g).bfloat16().cuda() if res else None
    rc = _lib.lib().gcb_debug_gemm_nt_bf16(A.data_ptr(), W.data_ptr(), Cm.data_ptr(), rows, N, K,
                                           None if b is None else b.data_ptr(), None if rp is None else rp.data_ptr(),
                                           0, None if R is None else R.data_ptr(), ACT[act],
                                           torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    if b is not None:
        w = (rp[1:] - rp[:-1]).float()[:, None] if rp is not None else 1.0
        ref = ref + w * b
    if R is not None:
        ref = ref + R.float()
    if act == "softplus":
        ref = F.softplus(ref)
    elif act == "relu":
        ref = F.relu(ref)
    elif act == "tanh":
        ref = torch.tanh(ref)
    err = (Cm.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-6)
    assert torch.isfinite(Cm.float()).all()
    assert float(err) < 2 ** -7, float(err)


@pytest.mark.parametrize("K,N", [(128, 256), (256, 128), (64, 128), (128, 64)])
@pytest.mark.parametrize("rows", [1, 127, 128, 129, 1000, 40000])
def test_tc_gemm_plain(rows, K, N):
    _run(rows, N, K)


@pytest.mark.parametrize("K,N", [(256, 512), (512, 256), (192, 768)])
@pytest.mark.parametrize("rows", [1, 129, 3000, 40000])
def test_tc_gemm_streamed_weights(rows, K, N):
    """Wide models (embedding_dim 256): weights do not fit shared memory and stream through the TMA ring."""
    _run(rows, N, K)
    if rows == 3000:
        _run(rows, N, K, bias=True)
        _run(rows, N, K, bias=True, deg=True, res=True, act="softplus")


@pytest.mark.parametrize("K,N", [(128, 256), (256, 128)])
def test_tc_gemm_epilogues(K, N):
    _run(5000, N, K, bias=True)
    _run(5000, N, K, bias=True, deg=True, res=True, act="softplus")
    _run(333, N, K, bias=True, res=True, act="tanh")
    _run(70000, N, K, bias=True, deg=True, res=True, act="relu")


def test_tc_gemm_is_deterministic_and_matches_cuda_core_path(monkeypatch):
    import os
    g = torch.Generator().manual_seed(1)
    A = torch.randn(3000, 128, generator=g).bfloat16().cuda()
    W = (torch.randn(256, 128, generator=g) / 11).bfloat16().cuda()
    outs = []
    for disable in ("0", "0", "1"):
        os.environ["GCB_DISABLE_TC"] = disable
        Cm = torch.empty(3000, 256, dtype=torch.bfloat16, device="cuda")
        assert _lib.lib().gcb_debug_gemm_nt_bf16(A.data_ptr(), W.data_ptr(), Cm.data_ptr(), 3000, 256, 128, None, None,
                                                 0, None, -1, torch.cuda.current_stream().cuda_stream) == 0
        torch.cuda.synchronize()
        outs.append(Cm.float())
    os.environ["GCB_DISABLE_TC"] = "0"
    assert torch.equal(outs[0], outs[1])
    assert float((outs[0] - outs[2]).abs().max() / outs[2].abs().max()) < 2 ** -7


@pytest.mark.parametrize("Nn,Kk", [(128, 256), (256, 128)])
@pytest.mark.parametrize("rows", [128, 129, 1000, 50000, 200000])
def test_tc_gemm_tn_weight_grad(rows, Nn, Kk):
    g = torch.Generator().manual_seed(rows)
    A = (torch.randn(rows, Nn, generator=g) * 0.3).bfloat16().cuda()
    B = torch.randn(rows, Kk, generator=g).bfloat16().cuda()
    d = torch.randint(0, 4, (rows,), generator=g)
    rp = torch.cat([torch.zeros(1, dtype=torch.long), d.cumsum(0)]).int().cuda()
    partial = torch.empty(148 * (Nn * Kk + Nn), device="cuda")
    outs = []
    for _ in range(2):
        dW = torch.full((Nn, Kk), float("nan"), device="cuda")
        db = torch.full((Nn,), float("nan"), device="cuda")
        rc = _lib.lib().gcb_debug_gemm_tn_bf16(A.data_ptr(), B.data_ptr(), rows, Nn, Kk, partial.data_ptr(), dW.data_ptr(),
                                               db.data_ptr(), rp.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        outs.append((dW.clone(), db.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), "must be deterministic"
    ref = A.double().t() @ B.double()
    refb = (A.double() * d.double().cuda()[:, None]).sum(0)
    assert float((dW.double() - ref).abs().max() / ref.abs().max()) < 1e-4
    # bias sums cover the first 128 columns only (all of dZA; the dP half of [dP|dQ]); the rest is zero
    assert float((db[:128].double() - refb[:128]).abs().max() / refb[:128].abs().max()) < 1e-4
    assert float(db[128:].abs().max()) == 0.0 if Nn > 128 else True


@pytest.mark.parametrize("Nn,Kk", [(256, 512), (512, 256)])
@pytest.mark.parametrize("rows", [128, 1000, 70001])
def test_tc_gemm_tn_weight_grad_wide(rows, Nn, Kk):
    """embedding_dim 256: dW is [256 x 512] / [512 x 256] -- the column-blocked tcgen05 kernel ([128 x 256] blocks, row ranges
    x blocks CTAs); bias sums over the first min(Nn, Kk) columns of A (all of dZA, the dP half of [dP|dQ])."""
    g = torch.Generator().manual_seed(rows + Nn)
    A = (torch.randn(rows, Nn, generator=g) * 0.3).bfloat16().cuda()
    B = torch.randn(rows, Kk, generator=g).bfloat16().cuda()
    d = torch.randint(0, 4, (rows,), generator=g)
    rp = torch.cat([torch.zeros(1, dtype=torch.long), d.cumsum(0)]).int().cuda()
    partial = torch.empty(148 * (Nn * Kk + Nn), device="cuda")
    lib = _lib.lib()
    lib.gcb_timing_enable(1)
    outs = []
    for _ in range(2):
        dW = torch.full((Nn, Kk), float("nan"), device="cuda")
        db = torch.full((Nn,), float("nan"), device="cuda")
        rc = lib.gcb_debug_gemm_tn_bf16(A.data_ptr(), B.data_ptr(), rows, Nn, Kk, partial.data_ptr(), dW.data_ptr(),
                                        db.data_ptr(), rp.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        outs.append((dW.clone(), db.clone()))
    import ctypes as C
    buf = C.create_string_buffer(1 << 16)
    lib.gcb_timing_report(buf, len(buf))
    lib.gcb_timing_enable(0)
    assert b"gemm_tn_tc_blocked" in buf.value, "the wide shapes must run on the tcgen05 kernel"
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), "must be deterministic"
    ref = A.double().t() @ B.double()
    refb = (A.double() * d.double().cuda()[:, None]).sum(0)
    assert float((dW.double() - ref).abs().max() / ref.abs().max()) < 1e-4
    nb = min(Nn, Kk)
    assert float((db[:nb].double() - refb[:nb]).abs().max() / refb[:nb].abs().max()) < 1e-4
    assert Nn == nb or float(db[nb:].abs().max()) == 0.0
