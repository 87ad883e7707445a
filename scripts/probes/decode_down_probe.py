import os, sys
sys.path.insert(0, "/root/repo")
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
I=14336
for rows in (4096, 3840, 3584, 2048, 1920):
    ws=[((torch.rand(rows, I, device=dev, generator=g)*2-1)/I**0.5).bfloat16() for _ in range(4)]
    for B in (1, 64):
        act=torch.randn(B, I, device=dev, generator=g).bfloat16()
        i=[0]
        def f():
            i[0]+=1; ops.linear_forward(act, ws[i[0]%4])
        t=timeit(f)
        print(f"rows {rows} B={B}: {t:.1f} us  {rows*I*2/t/1e3:.0f} GB/s  (clusters of 8: {rows//128})", flush=True)
