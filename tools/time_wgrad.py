"""Time pnr_wgrad (csrc/wgrad_tc05.cu) against the 3xTF32 library GEMM it replaces:
python tools/time_wgrad.py [S] [No] [Ni].  Default: 393216 samples (2048 rays x 192), 256 x 256."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from panopticnerf_b200.lib.train.mlp_backward import wgrad, _pow2_scale


def matmul_3xtf32(a, b, trans_a=False):   # what the training path used before: three TF32 library GEMMs on split operands
    def parts(t):
        hi = (t.contiguous().view(torch.int32) & -8192).view(torch.float32)
        return hi, t - hi
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    ah, al = parts(a); bh, bl = parts(b)
    if trans_a:
        ah, al = ah.t(), al.t()
    y = ah @ bh + (al @ bh + ah @ bl)
    torch.backends.cuda.matmul.allow_tf32 = old
    return y


DEV = "cuda:0"
S_ = int(sys.argv[1]) if len(sys.argv) > 1 else 393216
No = int(sys.argv[2]) if len(sys.argv) > 2 else 256
Ni = int(sys.argv[3]) if len(sys.argv) > 3 else 256
g = torch.Generator().manual_seed(0)
dz = (torch.randn(S_, No, generator=g) * 1e-6).to(DEV)
x = torch.relu(torch.randn(S_, Ni, generator=g)).to(DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, n=10):
    ts = []
    for i in range(n + 3):
        flush.zero_()                      # operands larger than L2 anyway; keep the partials cold too
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


t_k = timed(lambda: wgrad(dz, x))
sc = _pow2_scale(dz)
t_f = timed(lambda: wgrad(dz, x, precision="fp16x3", scale=sc))
t_s = timed(lambda: _pow2_scale(dz))
t_l = timed(lambda: (matmul_3xtf32(dz, x, trans_a=True), dz.sum(0)))
ref = dz.double().t() @ x.double()
e_k = float((wgrad(dz, x)[0].double() - ref).abs().max() / ref.pow(2).mean().sqrt())
e_f = float((wgrad(dz, x, precision="fp16x3", scale=sc)[0].double() - ref).abs().max() / ref.pow(2).mean().sqrt())
e_l = float((matmul_3xtf32(dz, x, trans_a=True).double() - ref).abs().max() / ref.pow(2).mean().sqrt())
byts = 4.0 * S_ * (No + Ni)
print(f"wgrad S={S_} No={No} Ni={Ni}: fp16x3 {t_f:.3f} ms err/rms {e_f:.1e} (+ scale reduction {t_s:.3f} ms) | bf16x3: pnr_wgrad {t_k:.3f} ms = {byts / t_k / 1e9:.2f} TB/s of operand reads, "
      f"{2.0 * S_ * No * Ni / t_k / 1e9:.0f} TFLOP/s algorithmic, err/rms {e_k:.1e} | 3xTF32 library GEMM + sum {t_l:.3f} ms, err/rms {e_l:.1e}")
