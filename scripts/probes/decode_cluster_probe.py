"""Same per-CTA work and CTA count as the 11B down projection (256 CTAs x 28 k-blocks) but WITHOUT the cluster-of-8 split-K:
a [32768, 1792] weight (256 row blocks, splits = 1).  If this streams much faster, cluster scheduling (15 co-resident clusters
of 8 = 120 SMs) is what holds the down projection back, not the memory system."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from llama32_b200 import ops
dev="cuda"; g=torch.Generator(device=dev).manual_seed(0)
def timeit(fn, iters=100, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters*1e3
for rows, K in ((32768, 1792), (4096, 14336), (37888, 1792), (18944, 3584)):
    ws=[((torch.rand(rows, K, device=dev, generator=g)*2-1)/K**0.5).bfloat16() for _ in range(4)]
    for B in (1, 64):
        act=torch.randn(B, K, device=dev, generator=g).bfloat16()
        i=[0]
        def f():
            i[0]+=1; ops.linear_forward(act, ws[i[0]%4])
        t=timeit(f)
        print(f"W[{rows},{K}] B={B}: {t:.1f} us  {rows*K*2/t/1e3:.0f} GB/s  row blocks {rows//128}", flush=True)
