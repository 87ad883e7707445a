"""A/B of the RMSNorm-backward kernel shapes (L32_RMSBWD_VARIANT) + numerics against torch fp32.  One process per variant."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) == 1:
    for v in ("0", "1", "2"):
        subprocess.run([sys.executable, __file__, v], env={**os.environ, "L32_RMSBWD_VARIANT": v})
    sys.exit(0)
import torch
from llama32_b200 import ops
dev = "cuda"
for T, H in ((8192, 4096), (8192, 8192), (1000, 2048), (333, 4096)):
    nb = 6
    g = torch.Generator(device=dev).manual_seed(0)
    dys = [torch.randn(T, H, device=dev, generator=g).bfloat16() for _ in range(nb)]
    hs = [torch.randn(T, H, device=dev, generator=g).bfloat16() for _ in range(nb)]
    gamma = (1 + 0.1 * torch.randn(H, device=dev, generator=g)).bfloat16()
    hf = hs[0].float()
    rms = torch.sqrt(hf.pow(2).mean(-1) + 1e-5)
    add = torch.randn(T, H, device=dev, generator=g).bfloat16()
    dx, dw = ops.rmsnorm_backward(dys[0], hs[0], gamma, rms)
    dxa, _ = ops.rmsnorm_backward(dys[0], hs[0], gamma, rms, addend=add)
    rstd = (1 / rms)[:, None]
    xhat = hf * rstd
    wdy = dys[0].float() * gamma.float()
    ref = rstd * (wdy - xhat * (wdy * xhat).mean(-1, keepdim=True))
    refw = (dys[0].float() * xhat).sum(0)
    e1 = ((dx.float() - ref).norm() / ref.norm()).item()
    e2 = ((dw.float() - refw).norm() / refw.norm()).item()
    e3 = ((dxa.float() - (ref + add.float())).norm() / (ref + add.float()).norm()).item()
    i = [0]
    def f():
        i[0] += 1
        ops.rmsnorm_backward(dys[i[0] % nb], hs[(i[0] + 1) % nb], gamma, rms)
    for _ in range(10): f()
    torch.cuda.synchronize()
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): f()
    e1_.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1_) * 10
    print(f"variant {sys.argv[1]} {T}x{H}: {us:.2f} us  {3 * T * H * 2 / us / 1e3:.0f} GB/s  rel dx {e1:.2e} dw {e2:.2e} dx+add {e3:.2e}", flush=True)
