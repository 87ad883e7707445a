"""Bring-up probe for the GPU box: runs one named case and prints error statistics against torch fp32.
Not a test (tests/ holds the pytest suite); used through `scripts/gpu_probe.sh` with one process per case so a
trapped kernel cannot poison the following cases."""
import sys, time
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from llama32_b200 import ops  # noqa: E402


def stats(name, got, ref):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    rel = (got - ref).norm() / ref.norm().clamp_min(1e-30)
    print(f"  {name}: rel_l2={rel.item():.3e} max_abs={err.max().item():.3e} max_ref={ref.abs().max().item():.3e} "
          f"nan={torch.isnan(got).sum().item()}", flush=True)
    return rel.item()


def timeit(fn, iters=20, warm=3):
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


def case_gemm(m, n, k, cta_group, a_mn=False, b_mn=False, dtype=torch.bfloat16, two_phase=False):
    torch.manual_seed(0)
    dev = "cuda"
    a = torch.randn(m, k, device=dev).to(dtype)
    b = torch.randn(n, k, device=dev).to(dtype)
    ref = a.float() @ b.float().t()
    a_in = a.t().contiguous() if a_mn else a
    b_in = b.t().contiguous() if b_mn else b
    kw = {}
    if two_phase:
        a1 = torch.randn(m, k, device=dev).to(dtype)
        b1 = torch.randn(n, k, device=dev).to(dtype)
        ref = ref + a1.float() @ b1.float().t()
        kw = dict(a1=a1.t().contiguous() if a_mn else a1, b1=b1.t().contiguous() if b_mn else b1)
    d = ops.gemm(a_in, b_in, a_mn_major=a_mn, b_mn_major=b_mn, cta_group=cta_group, **kw)
    torch.cuda.synchronize()
    print(f"gemm m={m} n={n} k={k} cta_group={cta_group} a_mn={a_mn} b_mn={b_mn} two_phase={two_phase} {dtype}")
    r = stats("D", d, ref)
    if r > 2e-2:
        # locate the damage: per 32x32 block error map summary
        e = (d.float() - ref).abs()
        rows = e.amax(dim=1)
        cols = e.amax(dim=0)
        print("   bad rows (first 16):", (rows > 0.5).nonzero().flatten()[:16].tolist(), "count", int((rows > 0.5).sum()))
        print("   bad cols (first 16):", (cols > 0.5).nonzero().flatten()[:16].tolist(), "count", int((cols > 0.5).sum()))
        print("   d[0,:8]", d[0, :8].float().tolist())
        print("   r[0,:8]", ref[0, :8].tolist())


def case_rmsnorm():
    torch.manual_seed(0)
    for dtype in (torch.bfloat16, torch.float16):
        for rows, c in ((7, 256), (33, 4096), (5, 8192), (3, 1000), (4, 250)):
            x = torch.randn(rows, c, device="cuda").to(dtype)
            r = torch.randn(rows, c, device="cuda").to(dtype)
            w = (1 + 0.1 * torch.randn(c, device="cuda")).to(dtype)
            for res in (None, r):
                y, rms, h = ops.add_rmsnorm_forward(x, w, res, 1e-5, want_h=True)
                hf = x.float() + (0 if res is None else res.float())
                var = hf.pow(2).mean(-1, keepdim=True) + 1e-5
                ref = hf * torch.rsqrt(var) * w.float()
                print(f"rmsnorm {dtype} rows={rows} C={c} residual={res is not None}")
                stats("y", y, ref)
                stats("rms", rms, var.sqrt().flatten())
                if h is not None:
                    stats("h", h, hf)
                g = torch.randn(rows, c, device="cuda").to(dtype)
                hh = hf.to(dtype)
                dx, dw = ops.rmsnorm_backward(g, hh, w, rms)
                h32 = hh.float().requires_grad_(True)
                w32 = w.float().requires_grad_(True)
                yy = h32 * torch.rsqrt(h32.pow(2).mean(-1, keepdim=True) + 1e-5) * w32
                yy.backward(g.float())
                stats("dx", dx, h32.grad)
                stats("dw", dw, w32.grad)


def case_swiglu(tokens, hidden, inter, cache):
    torch.manual_seed(0)
    dt = torch.bfloat16
    x = torch.randn(tokens, hidden, device="cuda").to(dt)
    wg = ((torch.rand(inter, hidden, device="cuda") * 2 - 1) / hidden ** 0.5).to(dt)
    wu = ((torch.rand(inter, hidden, device="cuda") * 2 - 1) / hidden ** 0.5).to(dt)
    wd = ((torch.rand(hidden, inter, device="cuda") * 2 - 1) / inter ** 0.5).to(dt)
    act, g, u = ops.swiglu_forward(x, wg, wu, want_cache=cache)
    gr = x.float() @ wg.float().t()
    ur = x.float() @ wu.float().t()
    ar = torch.nn.functional.silu(gr) * ur
    print(f"swiglu tokens={tokens} hidden={hidden} inter={inter} cache={cache}")
    stats("act", act, ar)
    if cache:
        stats("gate", g, gr)
        stats("up", u, ur)
    y, _, _ = ops.ffn_forward(x, wg, wu, wd)
    stats("ffn_y", y, ar @ wd.float().t())


def case_ffn_bwd(tokens, hidden, inter):
    torch.manual_seed(0)
    dt = torch.bfloat16
    x = torch.randn(tokens, hidden, device="cuda").to(dt)
    wg = ((torch.rand(inter, hidden, device="cuda") * 2 - 1) / hidden ** 0.5).to(dt)
    wu = ((torch.rand(inter, hidden, device="cuda") * 2 - 1) / hidden ** 0.5).to(dt)
    wd = ((torch.rand(hidden, inter, device="cuda") * 2 - 1) / inter ** 0.5).to(dt)
    dy = torch.randn(tokens, hidden, device="cuda").to(dt)
    y, g, u = ops.ffn_forward(x, wg, wu, wd, want_cache=True)
    dx, dwg, dwu, dwd, _, _ = ops.ffn_backward(dy, x, wg, wu, wd, g, u)
    xs, wgs, wus, wds = (t.float().requires_grad_(True) for t in (x, wg, wu, wd))
    yr = (torch.nn.functional.silu(xs @ wgs.t()) * (xs @ wus.t())) @ wds.t()
    yr.backward(dy.float())
    print(f"ffn_bwd tokens={tokens} hidden={hidden} inter={inter}")
    stats("y", y, yr)
    stats("dx", dx, xs.grad)
    stats("dw_gate", dwg, wgs.grad)
    stats("dw_up", dwu, wus.grad)
    stats("dw_down", dwd, wds.grad)


def case_perf(hidden, inter, tokens):
    dt = torch.bfloat16
    x = torch.randn(tokens, hidden, device="cuda").to(dt)
    wg = ((torch.rand(inter, hidden, device="cuda") * 2 - 1) / hidden ** 0.5).to(dt)
    wu = ((torch.rand(inter, hidden, device="cuda") * 2 - 1) / hidden ** 0.5).to(dt)
    wd = ((torch.rand(hidden, inter, device="cuda") * 2 - 1) / inter ** 0.5).to(dt)
    t_sw = timeit(lambda: ops.swiglu_forward(x, wg, wu))
    act, _, _ = ops.swiglu_forward(x, wg, wu)
    t_dn = timeit(lambda: ops.linear_forward(act, wd))
    t_ffn = timeit(lambda: ops.ffn_forward(x, wg, wu, wd))
    fl_sw, fl_dn = 4.0 * tokens * hidden * inter, 2.0 * tokens * hidden * inter
    print(f"perf H={hidden} I={inter} T={tokens}: swiglu {t_sw:.3f} ms ({fl_sw / t_sw / 1e9:.1f} TF/s)  down {t_dn:.3f} ms "
          f"({fl_dn / t_dn / 1e9:.1f} TF/s)  ffn {t_ffn:.3f} ms ({(fl_sw + fl_dn) / t_ffn / 1e9:.1f} TF/s, "
          f"{tokens / t_ffn * 1e3:.0f} tok/s)")
    # cuBLAS comparison (what the reference's live path runs on a GPU)
    t_cb = timeit(lambda: torch.nn.functional.linear(torch.nn.functional.silu(torch.nn.functional.linear(x, wg)) *
                                                     torch.nn.functional.linear(x, wu), wd))
    print(f"   torch/cuBLAS unfused ffn {t_cb:.3f} ms ({(fl_sw + fl_dn) / t_cb / 1e9:.1f} TF/s)")
    for rows, c in ((tokens, hidden),):
        r = torch.randn(rows, c, device="cuda").to(dt)
        w = torch.ones(c, device="cuda", dtype=dt)
        t = timeit(lambda: ops.add_rmsnorm_forward(x, w, r, 1e-5, want_rms=False), iters=50)
        print(f"   add_rmsnorm {rows}x{c}: {t * 1e3:.1f} us  {3 * rows * c * 2 / t / 1e6:.0f} GB/s")


if __name__ == "__main__":
    case = sys.argv[1]
    args = [int(v) for v in sys.argv[2:]]
    t0 = time.time()
    if case == "rmsnorm":
        case_rmsnorm()
    elif case == "gemm":
        m, n, k, cg, amn, bmn, tp = (args + [0, 0, 0])[:7]
        case_gemm(m, n, k, cg, bool(amn), bool(bmn), two_phase=bool(tp))
    elif case == "gemm16":
        m, n, k, cg = args[:4]
        case_gemm(m, n, k, cg, dtype=torch.float16)
    elif case == "swiglu":
        case_swiglu(args[0], args[1], args[2], bool(args[3]))
    elif case == "ffn_bwd":
        case_ffn_bwd(*args[:3])
    elif case == "perf":
        case_perf(*args[:3])
    print(f"[{case} {args}] done in {time.time() - t0:.1f}s", flush=True)
