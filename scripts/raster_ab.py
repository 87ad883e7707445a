import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llama32_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from perf_suite import timeit, weights
H, I, T = 4096, 14336, 8192
wg, wu, wd = weights(H, I)
x = torch.randn(T, H, device="cuda").bfloat16()
act = torch.randn(T, I, device="cuda").bfloat16()
for rep in range(2):
    for rg in (0, 16, 32):
        if rg: os.environ["L32_RASTER_GROUP"] = str(rg)
        else: os.environ.pop("L32_RASTER_GROUP", None)
        t1 = timeit(lambda: ops.swiglu_forward(x, wg, wu), iters=10)
        t2 = timeit(lambda: ops.linear_forward(act, wd), iters=10)
        print(f"raster={rg or 'default(8)'}: swiglu {t1*1e3:.0f} us ({4.0*T*H*I/t1/1e9:.0f} TF/s)  down {t2*1e3:.0f} us ({2.0*T*H*I/t2/1e9:.0f} TF/s)", flush=True)
        torch.cuda.synchronize(); import time; time.sleep(0.5)
