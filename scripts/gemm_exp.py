"""GEMM tuning experiments on the GPU box (timing only; correctness lives in tests/)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llama32_b200 import ops


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    dt = torch.bfloat16
    T, H, I = 8192, 4096, 14336
    x = torch.randn(T, H, device="cuda").to(dt)
    wg = (torch.randn(I, H, device="cuda") / 64).to(dt)
    wu = (torch.randn(I, H, device="cuda") / 64).to(dt)
    wd = (torch.randn(H, I, device="cuda") / 120).to(dt)
    act = torch.randn(T, I, device="cuda").to(dt)
    if mode == "one":
        # a single launch of each headline kernel for ncu
        ops.swiglu_forward(x, wg, wu)
        ops.linear_forward(act, wd)
        torch.cuda.synchronize()
        return
    fl = 2.0 * T * H * I
    for cg in (1, 2):
        for mc in (0,):
            t = timeit(lambda: ops.gemm(act, wd, cta_group=cg, max_ctas=mc))
            print(f"down-shape gemm cta_group={cg} max_ctas={mc}: {t:.3f} ms {fl / t / 1e9:.1f} TF/s", flush=True)
    t = timeit(lambda: torch.nn.functional.linear(act, wd))
    print(f"cuBLAS down-shape: {t:.3f} ms {fl / t / 1e9:.1f} TF/s")
    t = timeit(lambda: torch.nn.functional.linear(x, wg))
    print(f"cuBLAS gate-shape: {t:.3f} ms {fl / t / 1e9:.1f} TF/s")
    # small-K / compute-only scaling: is it the main loop or the tile turnaround?
    for k in (512, 2048, 8192):
        a = torch.randn(8192, k, device="cuda").to(dt)
        b = torch.randn(4096, k, device="cuda").to(dt)
        t = timeit(lambda: ops.gemm(a, b, cta_group=2))
        print(f"gemm 8192x4096x{k} cg2: {t:.3f} ms {2.0 * 8192 * 4096 * k / t / 1e9:.1f} TF/s", flush=True)
    # fewer CTAs: per-SM rate when L2 is not contended
    for mc in (2, 16, 64, 148):
        a = torch.randn(256 * (mc // 2), 4096, device="cuda").to(dt)
        b = torch.randn(256, 4096, device="cuda").to(dt)
        t = timeit(lambda: ops.gemm(a, b, cta_group=2, max_ctas=mc))
        print(f"one tile per pair, {mc} CTAs, K=4096: {t * 1e3:.1f} us  per-pair {2.0 * 256 * 256 * 4096 / t / 1e9:.2f} TF/s "
              f"(peak/pair ~ {2250 / 74:.1f})", flush=True)


if __name__ == "__main__":
    main()
