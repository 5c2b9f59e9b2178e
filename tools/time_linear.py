"""Time pnr_linear (csrc/linear_tc05.cu) against the 3xTF32 library GEMM it replaces:
python tools/time_linear.py [S] [K] [N].  Default: 393216 samples (2048 rays x 192), 256 -> 256 forward, then the
283 -> 128 view layer with ReLU and its input gradient."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from panopticnerf_b200.lib.train.mlp_backward import linear3x, _pow2_scale

DEV = "cuda:0"
S_ = int(sys.argv[1]) if len(sys.argv) > 1 else 393216
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def tf32x3(a, b):   # what the training path used before: three TF32 library GEMMs on split operands
    def parts(t):
        hi = (t.contiguous().view(torch.int32) & -8192).view(torch.float32)
        return hi, t - hi
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    ah, al = parts(a); bh, bl = parts(b)
    y = ah @ bh + (al @ bh + ah @ bl)
    torch.backends.cuda.matmul.allow_tf32 = old
    return y


def timed(fn, n=10):
    ts = []
    for i in range(n + 3):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def case(K, N, relu, transposed, gscale, prec):
    g = torch.Generator().manual_seed(K + N)
    x = (torch.randn(S_, N if transposed else K, generator=g) * gscale).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = None if transposed else (torch.randn(N, generator=g) * 0.1).to(DEV)
    sc = _pow2_scale(x) if (transposed and prec == "fp16x3") else None
    pad = (-(K if transposed else N)) % 8 if transposed else 0      # the training path pads gradient rows to 32 bytes
    f = lambda: linear3x(x, w, b, relu=relu, transposed=transposed, precision=prec, scale=sc,
                         out_cols=(K + pad) if transposed else 0)[:, :(K if transposed else N)]
    lib = lambda: (tf32x3(x, w) if transposed else tf32x3(x, w.t()) + b)
    t_k, t_l = timed(f), timed(lib)
    ref = (x.double() @ w.double()) if transposed else (x.double() @ w.double().t() + b.double())
    if relu:
        ref = ref.clamp(min=0)
    e = float((f().double() - ref).abs().max() / ref.pow(2).mean().sqrt())
    kin, nout = (N, K) if transposed else (K, N)
    byts = 4.0 * S_ * (kin + nout)
    print(f"linear S={S_} {kin}->{nout}{' relu' if relu else ''}{' (input gradient)' if transposed else ''} {prec}: pnr_linear {t_k:.3f} ms = "
          f"{byts / t_k / 1e9:.2f} TB/s of x + y, err/rms {e:.1e} | 3xTF32 library GEMM {t_l:.3f} ms")


if len(sys.argv) > 3:
    case(int(sys.argv[2]), int(sys.argv[3]), False, False, 1.0, "fp16x3")
else:
    case(256, 256, False, False, 1.0, "fp16x3")
    case(283, 128, True, False, 1.0, "fp16x3")
    case(283, 128, False, True, 3e-7, "fp16x3")
