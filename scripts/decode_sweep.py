"""Knob sweep of the small-M weight-streaming kernels: python scripts/decode_sweep.py  (one subprocess per setting)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) == 1:
    settings = [{}, {"L32_DECODE_THREADS": "128"}, {"L32_DECODE_XSTAGES": "4", "L32_DECODE_STAGES": "5"}, {"L32_DECODE_DEBUG_NOX": "1"}]
    for st in settings:
        subprocess.run([sys.executable, __file__, "run", repr(st)], env={**os.environ, **st})
    sys.exit(0)
import torch
from llama32_b200 import ops
dev = "cuda"
H, I = 4096, 14336
g = torch.Generator(device=dev).manual_seed(0)
uni = lambda r, c: ((torch.rand(r, c, device=dev, generator=g) * 2 - 1) / c ** 0.5).bfloat16()
ws = [(uni(I, H), uni(I, H), uni(H, I)) for _ in range(3)]
def timeit(fn, iters=100, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
out = []
for B in (1, 16, 32, 48, 64, 128):
    x = torch.randn(B, H, device=dev, generator=g).bfloat16()
    act = torch.randn(B, I, device=dev, generator=g).bfloat16()
    i = [0]
    def nxt():
        i[0] += 1
        return ws[i[0] % 3]
    gamma = torch.ones(H, device=dev, dtype=torch.bfloat16)
    r = torch.randn(B, H, device=dev, generator=g).bfloat16()
    t_blk = timeit(lambda: ops.ffn_forward(ops.add_rmsnorm_forward(x, gamma, r, 1e-5, want_rms=False)[0], *nxt()))
    t_ffn = timeit(lambda: ops.ffn_forward(x, *nxt()))
    t_gu = timeit(lambda: ops.swiglu_forward(x, *nxt()[:2]))
    t_dn = timeit(lambda: ops.linear_forward(act, nxt()[2]))
    out.append(f"B={B}: norm+ffn {t_blk:.1f} ffn {t_ffn:.1f} gu {t_gu:.1f} ({2*H*I*2/t_gu/1e3:.0f} GB/s) dn {t_dn:.1f} ({H*I*2/t_dn/1e3:.0f} GB/s)")
print(sys.argv[2], " | ".join(out), flush=True)
