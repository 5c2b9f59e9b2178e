"""VERDICT r1, "next" item 9 (cheap CPU experiment): can the two cross terms of the 3-pass split run at FP8 rate?

fp16x3 computes  x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo  with every operand a 11-bit fp16 value; the two cross
terms are ~2^-11 of the product, so if they tolerated 4-bit operands (kind::f8f6f4, twice the fp16 MMA rate) the
scheme would cost 2 pass-equivalents instead of 3 (bound of the algorithmic roofline fraction 0.5 instead of 0.375).
kind::f8f6f4 takes BOTH operands in <= 8 bits, so x_hi and w_hi have to be re-rounded to e4m3 / e5m2 for those
terms as well.  This script runs the cfg2 network (8x256, default-init weights, 4000 samples) with
    A  fp16x3                         (what the kernel does)
    B  hi*hi in fp16, cross terms with e4m3 operands (per-tensor power-of-two scale so nothing saturates)
    C  the same with e5m2 operands
    D  hi*hi only                     (the 1-pass fast mode)
and prints the max error of sigma / rgb relative to the per-tensor RMS (the parity tolerance is 1e-4).
Accumulation is exact (float64) in every variant: only the operand rounding differs."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import panopticnerf_b200 as PN                        # noqa: E402
from oracle import reference_renderer as O            # noqa: E402
from panopticnerf_b200 import synthetic as S          # noqa: E402


def split16(x):
    hi = x.to(torch.float16).to(torch.float64)
    lo = (x - hi).to(torch.float16).to(torch.float64)
    return hi, lo


def q8(x, dtype):
    """Round to an 8-bit float with a per-tensor power-of-two scale that maps max|x| just below the format's max."""
    fmax = 448.0 if dtype == torch.float8_e4m3fn else 57344.0
    m = float(x.abs().max())
    if m == 0.0:
        return x
    scale = 2.0 ** torch.floor(torch.log2(torch.tensor(fmax / m))).item()
    return (x * scale).to(torch.float32).to(dtype).to(torch.float64) / scale


def matmul(x, w, mode):
    xh, xl = split16(x)
    wh, wl = split16(w)
    y = xh @ wh.T
    if mode == "fp16x3":
        y = y + xl @ wh.T + xh @ wl.T
    elif mode in ("e4m3", "e5m2"):
        dt = torch.float8_e4m3fn if mode == "e4m3" else torch.float8_e5m2
        y = y + q8(xl, dt) @ q8(wh, dt).T + q8(xh, dt) @ q8(wl, dt).T
    return y


def forward(net, pts, vd, mode):
    cfg = net.cfg if hasattr(net, "cfg") else None
    ex = O.embed(pts, 10).double()
    ed = O.embed(vd, 4).double()
    lin = lambda l, x: matmul(x, l.weight.detach().double(), mode) + l.bias.detach().double()
    h = ex
    D = len(net.pts_linears)
    for i, l in enumerate(net.pts_linears):
        h = torch.relu(lin(l, h))
        if i == D // 2:
            h = torch.cat([ex, h], -1)
    sigma = lin(net.alpha_linear, h)
    feat = lin(net.feature_linear, h)
    g = torch.relu(lin(net.views_linears[0], torch.cat([feat, ed], -1)))
    rgb = lin(net.rgb_linear, g)
    return torch.cat([rgb, sigma], -1)


def main():
    cfg = PN.make_cfg("cfg2")
    net = S.init_network_weights(O.make_network(cfg), seed=1)
    g = torch.Generator().manual_seed(9)
    pts = (torch.rand(4000, 3, generator=g) * 2 - 1) * 4
    vd = torch.nn.functional.normalize(torch.randn(4000, 3, generator=g), dim=-1)
    with torch.no_grad():
        ref = net(pts, vd).double()[:, :4]
        exact = forward(net, pts, vd, "exact64") if False else None
    rms = lambda t: float(torch.sqrt(torch.mean(t ** 2)))
    print(f"{'mode':28s} {'rgb max err / RMS':>20s} {'sigma max err / RMS':>20s}")
    for mode, label in (("fp16x3", "A fp16x3 (3 fp16 passes)"), ("e4m3", "B cross terms in e4m3"),
                        ("e5m2", "C cross terms in e5m2"), ("hi", "D hi*hi only (1 pass)")):
        with torch.no_grad():
            out = forward(net, pts, vd, mode)
        e_rgb = float((out[:, :3] - ref[:, :3]).abs().max()) / rms(ref[:, :3])
        e_sig = float((out[:, 3:] - ref[:, 3:]).abs().max()) / rms(ref[:, 3:])
        print(f"{label:28s} {e_rgb:20.3e} {e_sig:20.3e}")
    print("tolerance: 1e-4")


if __name__ == "__main__":
    main()
