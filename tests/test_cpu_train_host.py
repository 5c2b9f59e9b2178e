"""Host-side logic of the training path's GEMM wrappers (lib/train/mlp_backward.py): the power-of-two gradient scales
and the argument checks that run before any kernel - no GPU needed.  The kernels themselves are covered by
tests/test_gpu_backward.py against float64."""
import math

import pytest
import torch

from panopticnerf_b200 import _capi
from panopticnerf_b200.lib.train import mlp_backward as MB


def test_pow2_scale_is_a_power_of_two_that_brings_the_maximum_to_about_256():
    g = torch.Generator().manual_seed(0)
    for mag in (3e-7, 1.0, 4.0e4, 1e-20):
        t = torch.randn(50, 7, generator=g) * mag
        s = float(MB._pow2_scale(t))
        assert s == 2.0 ** round(math.log2(s))                        # exactly a power of two: scaling is exact
        assert 128.0 <= float(t.abs().max()) * s <= 512.0
    assert float(MB._pow2_scale(torch.zeros(3, 3))) == 1.0            # nothing to scale
    assert float(MB._pow2_scale(torch.tensor([[float("nan"), 1.0]]))) == 1.0
    assert float(MB._pow2_scale(torch.tensor([[float("inf"), 1.0]]))) == 1.0
    assert float(MB._pow2_scale(torch.tensor([[1e-38]]))) == 2.0 ** 100   # clamped: stays finite


def test_pow2_scales_equals_per_slice_scale():
    g = torch.Generator().manual_seed(1)
    t = torch.randn(5, 40, 16, generator=g) * torch.tensor([1e-6, 3e-3, 0.0, 7.0, 2e-9]).view(5, 1, 1)
    many = MB._pow2_scales(t)
    assert len(many) == 5 and all(m.shape == (1,) and m.is_contiguous() for m in many)
    for i in range(5):
        assert float(many[i]) == float(MB._pow2_scale(t[i]))


def test_rows_keeps_row_strided_views_and_copies_what_the_kernels_cannot_address():
    t = torch.arange(60, dtype=torch.float32).reshape(6, 10)
    # _rows refuses CPU tensors (no fallback) - the stride logic is exercised through a meta-free shim
    with pytest.raises(_capi.PnrError):
        MB._rows(t, "t")
    with pytest.raises(_capi.PnrError):
        MB.wgrad(t, t)
    with pytest.raises(_capi.PnrError):
        MB.linear3x(t, t)


def test_new_entry_points_validate_their_arguments_before_touching_the_device():
    """pnr_wgrad / pnr_linear refuse bad arguments with PNR_ERR_ARG and a message (no CUDA call is made on that path,
    so this runs without a GPU); the workspace queries are pure arithmetic."""
    L = _capi.lib()
    assert L.pnr_linear_workspace_bytes(256, 512) == 8 * 2 * 8 * 256 * 16      # 8 K-chunks of hi + lo images
    assert L.pnr_linear_workspace_bytes(300, 64) == 0 and L.pnr_linear_workspace_bytes(16, 600) == 0
    assert L.pnr_wgrad_workspace_bytes(300, 16) == 0 and L.pnr_wgrad_workspace_bytes(16, 0) == 0
    n = L.pnr_wgrad_workspace_bytes(256, 256)
    assert n > 0 and n % ((2 * 128 * 256 + 256) * 4) == 0                     # one partial product + bias row per SM
    prec = _capi.PREC["fp16x3"]
    rc = L.pnr_wgrad(None, 256, 256, None, 256, 256, 10, prec, None, None, 256, None, 0, None, 0, None)
    assert rc == -1 and b"required" in L.pnr_last_error()
    rc = L.pnr_wgrad(64, 256, 300, 64, 256, 256, 10, prec, None, 64, 256, None, 0, None, 0, None)   # (never dereferenced)
    assert rc == -1 and b"No = 300" in L.pnr_last_error()
    rc = L.pnr_wgrad(64, 100, 256, 64, 256, 256, 10, prec, None, 64, 256, None, 0, None, 0, None)
    assert rc == -1 and b"leading dimensions" in L.pnr_last_error()
    rc = L.pnr_wgrad(64, 256, 256, 64, 256, 256, 10, _capi.PREC["fp16"], None, 64, 256, None, 0, None, 0, None)
    assert rc == -1 and b"x3 precisions" in L.pnr_last_error()
    rc = L.pnr_linear(None, 256, 256, None, 256, 0, None, 256, 10, 0, prec, None, None, 256, None, 0, None)
    assert rc == -1 and b"required" in L.pnr_last_error()
    rc = L.pnr_linear(64, 256, 600, 64, 600, 0, None, 256, 10, 0, prec, None, 64, 256, None, 0, None)
    assert rc == -1 and b"K = 600" in L.pnr_last_error()
    rc = L.pnr_linear(64, 256, 256, 64, 256, 0, None, 256, 10, 0, prec, None, 64, 256, None, 0, None)
    assert rc == -1 and b"workspace" in L.pnr_last_error()


def test_operand_split_error_model_of_the_gradient_gemms():
    """What the 16-bit hi / lo operand parts of pnr_wgrad / pnr_linear can and cannot represent, with exact (float64)
    accumulation: hi.hi + lo.hi + hi.lo on ~1e-6 gradients.  fp16 parts need the power-of-two scale (unscaled, the
    gradients fall below fp16's subnormal spacing), and with it they are ~50x more accurate than bf16 parts; both are
    inside the path's 1e-4.  These are the levels the GPU tests' thresholds are set against."""
    g = torch.Generator().manual_seed(0)
    S_, No, Ni = 4096, 64, 96
    dz = torch.randn(S_, No, generator=g) * 1e-6 * (0.1 + torch.rand(1, No, generator=g) * 3.0)
    dz = dz * (torch.rand(S_, No, generator=g) < 0.6)
    x = torch.relu(torch.randn(S_, Ni, generator=g)) + 0.05 * torch.randn(S_, Ni, generator=g)
    ref = dz.double().t() @ x.double()

    def three_products(a, b, dt):
        ah = a.to(dt).float()
        al = (a - ah).to(dt).float()
        bh = b.to(dt).float()
        bl = (b - bh).to(dt).float()
        ah, al, bh, bl = (t.double() for t in (ah, al, bh, bl))
        return ah.t() @ bh + al.t() @ bh + ah.t() @ bl

    err = lambda y: float((y - ref).abs().max() / ref.pow(2).mean().sqrt())
    sc = float(MB._pow2_scale(dz))
    e_fp16 = err(three_products(dz * sc, x, torch.float16) / sc)
    e_bf16 = err(three_products(dz, x, torch.bfloat16))
    e_raw = err(three_products(dz, x, torch.float16))
    assert e_fp16 < 2e-6 and 5e-6 < e_bf16 < 6e-5 and e_fp16 < e_bf16 / 10
    assert e_raw > 1e-3                                  # unscaled fp16 parts: the gradients are mostly gone
