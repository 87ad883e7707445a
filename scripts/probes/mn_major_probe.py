"""Does an MN-major operand cost tensor-core throughput?  Same [M, N, K] problem with every combination of operand majors
(the transposes are consumed in place), at a tile count that divides the machine (no wave quantisation): M = 9472 = 37 x 256,
N = 4096 -> 592 tiles = 8 full waves of 74 clusters; K = 8192."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from llama32_b200 import ops
dt = torch.bfloat16
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
M, N, K = 9472, 4096, 8192
a_k = torch.randn(M, K, device="cuda").to(dt); a_mn = a_k.t().contiguous()      # [K, M]
b_k = torch.randn(N, K, device="cuda").to(dt); b_mn = b_k.t().contiguous()      # [K, N]
fl = 2.0 * M * N * K
ref = None
for am in (False, True):
    for bm in (False, True):
        a = a_mn if am else a_k
        b = b_mn if bm else b_k
        d = ops.gemm(a, b, a_mn_major=am, b_mn_major=bm)
        if ref is None: ref = d
        err = ((d.float() - ref.float()).norm() / ref.float().norm()).item()
        t = timeit(lambda: ops.gemm(a, b, a_mn_major=am, b_mn_major=bm))
        print(f"A {'MN' if am else 'K '}-major  B {'MN' if bm else 'K '}-major: {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s  (vs K/K result rel {err:.1e})", flush=True)
