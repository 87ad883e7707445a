"""Timing suite for the GPU box: decode, training step, 90B shape, raster sweep (numbers for DESIGN.md)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llama32_b200 import ops
import llama32_b200 as L


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


def weights(H, I, n=1):
    dt = torch.bfloat16
    out = []
    for _ in range(n):
        out.append(((torch.rand(I, H, device="cuda") * 2 - 1) / H ** 0.5).to(dt))
        out.append(((torch.rand(I, H, device="cuda") * 2 - 1) / H ** 0.5).to(dt))
        out.append(((torch.rand(H, I, device="cuda") * 2 - 1) / I ** 0.5).to(dt))
    return out


def decode(H, I, nsets=3, batches=(1, 4, 16, 32, 64, 128), knobs=({},)):
    # rotate weight sets so every step streams weights from HBM (3 x 352 MB > 126 MB L2)
    ws = weights(H, I, nsets)
    gamma = torch.ones(H, device="cuda", dtype=torch.bfloat16)
    for kn in knobs:
        for k in list(os.environ):
            if k.startswith("L32_DECODE_"):
                del os.environ[k]
        os.environ.update({k: str(v) for k, v in kn.items()})
        for B in batches:
            x = torch.randn(B, 1, H, device="cuda").bfloat16()
            r = torch.randn(B, 1, H, device="cuda").bfloat16()
            act = torch.randn(B, 1, I, device="cuda").bfloat16()
            state = {"i": 0}

            def nxt():
                i = state["i"] % nsets
                state["i"] += 1
                return ws[3 * i], ws[3 * i + 1], ws[3 * i + 2]

            def f():
                wg, wu, wd = nxt()
                ops.ffn_forward(x, wg, wu, wd)

            def g():
                wg, wu, wd = nxt()
                ops.swiglu_forward(x, wg, wu)

            def d():
                wg, wu, wd = nxt()
                ops.linear_forward(act, wd)

            def blk():
                wg, wu, wd = nxt()
                ops.ffn_forward(ops.add_rmsnorm_forward(x, gamma, r, 1e-5, want_rms=False)[0], wg, wu, wd)
            t, tg, td, tb = (timeit(fn, iters=60, warm=6) for fn in (f, g, d, blk))
            gb = 3.0 * H * I * 2 / 1e9
            print(f"decode H={H} I={I} B={B} {kn}: ffn {t * 1e3:.1f} us ({gb / t * 1e3:.0f} GB/s)  gate/up {tg * 1e3:.1f} us "
                  f"({gb * 2 / 3 / tg * 1e3:.0f} GB/s)  down {td * 1e3:.1f} us ({gb / 3 / td * 1e3:.0f} GB/s)  "
                  f"norm+ffn {tb * 1e3:.1f} us", flush=True)
    for k in list(os.environ):
        if k.startswith("L32_DECODE_"):
            del os.environ[k]


def train(H, I, T):
    dt = torch.bfloat16
    norm = L.LLAMARMSNorm(H, eps=1e-5).to("cuda", dt)
    ffn = L.FusedFeedforward(H, I).to("cuda", dt)
    x = torch.randn(T, H, device="cuda").to(dt)
    r = torch.randn(T, H, device="cuda").to(dt)
    dy = torch.randn(T, H, device="cuda").to(dt)

    def step():
        xx = x.detach().requires_grad_(True)
        y = ffn(norm(xx, residual=r))
        y.backward(dy)
    t = timeit(step, iters=10)
    fl = 18.0 * T * H * I
    print(f"train H={H} I={I} T={T}: fwd+bwd {t:.3f} ms  {fl / t / 1e9:.0f} TF/s  {T / t * 1e3:.0f} tok/s", flush=True)
    wg, wu, wd = ffn.swiglu.w_gate.detach(), ffn.swiglu.w_up.detach(), ffn.w_down.weight.detach()
    xn = x
    y, g, u = ops.ffn_forward(xn, wg, wu, wd, want_cache=True)
    t_f = timeit(lambda: ops.ffn_forward(xn, wg, wu, wd, want_cache=True), iters=10)
    t_b = timeit(lambda: ops.ffn_backward(dy, xn, wg, wu, wd, g, u), iters=10)
    print(f"   ffn fwd(+caches) {t_f:.3f} ms ({6.0 * T * H * I / t_f / 1e9:.0f} TF/s)   ffn bwd {t_b:.3f} ms ({12.0 * T * H * I / t_b / 1e9:.0f} TF/s)")
    # individual backward GEMMs
    dg = torch.randn(T, I, device="cuda").to(dt)
    t1 = timeit(lambda: ops.gemm(dy, wd, b_mn_major=True), iters=10)                      # d_act
    t2 = timeit(lambda: ops.gemm(dg, wg, b_mn_major=True, a1=dg, b1=wu), iters=10)         # dx two-phase
    t3 = timeit(lambda: ops.gemm(dg, xn, a_mn_major=True, b_mn_major=True), iters=10)      # wgrad
    f2 = 2.0 * T * H * I
    print(f"   d_act gemm {t1:.3f} ms ({f2 / t1 / 1e9:.0f} TF/s)  dx 2-phase {t2:.3f} ms ({2 * f2 / t2 / 1e9:.0f} TF/s)  wgrad {t3:.3f} ms ({f2 / t3 / 1e9:.0f} TF/s)")


def norm(H, T, nbuf=6):
    """Add-RMSNorm / RMSNorm backward HBM bandwidth; nbuf rotating buffer sets (> 126 MB L2) so nothing is L2-resident."""
    dt = torch.bfloat16
    xs = [torch.randn(T, H, device="cuda").to(dt) for _ in range(nbuf)]
    rs = [torch.randn(T, H, device="cuda").to(dt) for _ in range(nbuf)]
    dys = [torch.randn(T, H, device="cuda").to(dt) for _ in range(nbuf)]
    rms = torch.rand(T, device="cuda") + 0.5
    w = (1 + 0.1 * torch.randn(H, device="cuda")).to(dt)
    st = {"i": 0}

    def nx():
        st["i"] += 1
        return st["i"] % nbuf
    tb = timeit(lambda: ops.rmsnorm_backward(dys[nx()], xs[nx()], w, rms), iters=60, warm=6)
    tb0 = timeit(lambda: ops.rmsnorm_backward(dys[nx()], xs[nx()], w, rms, want_dweight=False), iters=60, warm=6)
    tf = timeit(lambda: ops.add_rmsnorm_forward(xs[nx()], w, rs[nx()], 1e-5, want_h=True), iters=60, warm=6)
    tf3 = timeit(lambda: ops.add_rmsnorm_forward(xs[nx()], w, rs[nx()], 1e-5, want_h=False, want_rms=False), iters=60, warm=6)
    tf2 = timeit(lambda: ops.add_rmsnorm_forward(xs[nx()], w, None, 1e-5, want_h=False, want_rms=False), iters=60, warm=6)
    b = T * H * 2
    print(f"norm H={H} T={T}: fwd(no res) {tf2 * 1e3:.1f} us {2 * b / tf2 / 1e6:.0f} GB/s | add-rmsnorm fwd {tf3 * 1e3:.1f} us "
          f"{3 * b / tf3 / 1e6:.0f} GB/s | +h {tf * 1e3:.1f} us {4 * b / tf / 1e6:.0f} GB/s | bwd {tb * 1e3:.1f} us {3 * b / tb / 1e6:.0f} GB/s (no dweight reduce: {tb0 * 1e3:.1f} us)",
          flush=True)


def prefill(H, I, T):
    wg, wu, wd = weights(H, I)
    x = torch.randn(T, H, device="cuda").bfloat16()
    act = torch.randn(T, I, device="cuda").bfloat16()
    for rg in (0, 2, 4, 8, 16, 32):
        if rg:
            os.environ["L32_RASTER_GROUP"] = str(rg)
        else:
            os.environ.pop("L32_RASTER_GROUP", None)
        t1 = timeit(lambda: ops.swiglu_forward(x, wg, wu), iters=10)
        t2 = timeit(lambda: ops.linear_forward(act, wd), iters=10)
        print(f"prefill H={H} I={I} T={T} raster={rg or 'default'}: swiglu {t1:.3f} ms ({4.0 * T * H * I / t1 / 1e9:.0f} TF/s) "
              f"down {t2:.3f} ms ({2.0 * T * H * I / t2 / 1e9:.0f} TF/s)", flush=True)
    os.environ.pop("L32_RASTER_GROUP", None)


def tp_shapes():
    """The per-rank GEMMs of the tensor-parallel path, run locally (no NVLink): isolates wave quantisation / short-K
    effects from the cost of the fused collectives."""
    T = 8192
    for name, H, I in (("11b", 4096, 14336), ("90b", 8192, 28672)):
        for p in (2, 4, 8):
            Il = I // p
            wg, wu, wd = weights(H, Il)
            x = torch.randn(T, H, device="cuda").bfloat16()
            act = torch.randn(T, Il, device="cuda").bfloat16()
            t1 = timeit(lambda: ops.swiglu_forward(x, wg, wu), iters=20)
            t2 = timeit(lambda: ops.linear_forward(act, wd), iters=20)
            print(f"tp-shape {name} p={p} I/p={Il}: gate/up {t1 * 1e3:.0f} us ({4.0 * T * H * Il / t1 / 1e9:.0f} TF/s)  "
                  f"down {t2 * 1e3:.0f} us ({2.0 * T * H * Il / t2 / 1e9:.0f} TF/s)", flush=True)


def tp_emul(world=8, H=4096, I=14336, T=8192, one_kernel=False):
    """Fused TP phases with all ranks emulated on one GPU (peer pointers = local buffers): cost of the mechanism itself
    (rotation, raster, flag waits, pulls / pushes through local memory) without NVLink."""
    from llama32_b200.tp import FusedTensorParallelBlock, TpRankBuffers
    dt = torch.bfloat16
    wg, wu, wd = weights(H, I)
    gamma = torch.ones(H, device="cuda", dtype=dt)
    bufs = TpRankBuffers.local_world(world, T, H, dt, "cuda")
    blocks = [FusedTensorParallelBlock(gamma, 1e-5, wg, wu, wd, b, one_kernel=one_kernel) for b in bufs]
    x = torch.randn(T, H, device="cuda").to(dt)
    r = torch.randn(T, H, device="cuda").to(dt)
    acc = [0.0, 0.0, 0.0, 0.0]
    iters = 10
    for it in range(iters + 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        for blk in blocks:
            lo, hi, _ = blk.rows_of(T)
            blk.phase_norm(x[lo:hi], r[lo:hi], T)
        ev[1].record()
        if one_kernel:
            for blk in blocks:
                blk.phase_ffn(T)
            ev[2].record()
        else:
            for blk in blocks:
                blk.phase_gate_up(T)
            ev[2].record()
            for blk in blocks:
                blk.phase_down(T)
        ev[3].record()
        for blk in blocks:
            blk.phase_reduce(T)
        ev[4].record()
        torch.cuda.synchronize()
        if it >= 2:
            for k in range(4):
                acc[k] += ev[k].elapsed_time(ev[k + 1])
    per = [a / iters / world * 1e3 for a in acc]
    print(f"tp-emul world={world} H={H} I={I} T={T} one_kernel={one_kernel}: per rank norm+signal {per[0]:.0f} us  gate/up+pull {per[1]:.0f} us  "
          f"down+push+signal {per[2]:.0f} us  reduce {per[3]:.0f} us", flush=True)


def ffn1k():
    """One-kernel feed-forward (EPI_FFN_TP, world = 1: no NVLink) against the two separate GEMM kernels."""
    dt = torch.bfloat16
    for T, H, I in ((8192, 4096, 1792), (32768, 4096, 1792)):
        wg, wu, wd = weights(H, I)
        x = torch.randn(T, H, device="cuda").to(dt)
        slot = torch.empty(T, H, device="cuda", dtype=dt)
        ready = torch.zeros(8, dtype=torch.int32, device="cuda")
        done = torch.zeros(8, dtype=torch.int32, device="cuda")
        act = torch.empty(T, I, device="cuda", dtype=dt)
        act_done = torch.empty((T + 255) // 256, dtype=torch.int32, device="cuda")

        def two():
            a, _, _ = ops.swiglu_forward(x, wg, wu)
            ops.linear_forward(a, wd)

        def one():
            ops.tp_ffn_forward_fused(x, [x.data_ptr()], ready, done, 1, 0, T, wg, wu, wd, [slot.data_ptr()], act=act, act_done=act_done)
        fl = 6.0 * T * H * I
        for env in ({}, {"L32_RASTER_GROUP": "4"}, {"L32_RASTER_GROUP": "32"}):
            for k in ("L32_FFN_PREFIX", "L32_RASTER_GROUP"):
                os.environ.pop(k, None)
            os.environ.update(env)
            t1 = timeit(one, iters=20)
            print(f"ffn1k T={T} H={H} I={I} {env}: one kernel {t1 * 1e3:.0f} us ({fl / t1 / 1e9:.0f} TF/s)", flush=True)
        for k in ("L32_FFN_PREFIX", "L32_RASTER_GROUP"):
            os.environ.pop(k, None)
        t2 = timeit(two, iters=20)
        print(f"ffn1k T={T} H={H} I={I}: two kernels {t2 * 1e3:.0f} us ({fl / t2 / 1e9:.0f} TF/s)", flush=True)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "decode_mid":
        decode(4096, 14336, batches=(24, 40, 48, 64, 80, 96))
    elif what == "decode":
        decode(4096, 14336)
        decode(8192, 28672, nsets=2)
    elif what == "decode_sweep":
        decode(4096, 14336, batches=(1, 16, 32, 64, 128), knobs=({}, {"L32_DECODE_ROTATE": 0}))
        decode(8192, 28672, nsets=2, batches=(1, 64, 128), knobs=({}, {"L32_DECODE_ROTATE": 0}))
    elif what == "train":
        train(4096, 14336, 8192)
    elif what == "tp_shapes":
        tp_shapes()
    elif what == "ffn1k":
        ffn1k()
    elif what == "tp_emul":
        tp_emul(8, one_kernel=False)
        tp_emul(8, T=32768, one_kernel=False)
        for env in ({}, {"L32_RASTER_GROUP": "32"}, {"L32_RASTER_GROUP": "16"}):
            for k in ("L32_FFN_PREFIX", "L32_SWIGLU_TILE_N", "L32_RASTER_GROUP"):
                os.environ.pop(k, None)
            os.environ.update(env)
            print(env)
            tp_emul(8, one_kernel=True)
            tp_emul(8, T=32768, one_kernel=True)
    elif what == "norm":
        norm(4096, 8192)
        norm(8192, 8192)
        norm(4096, 64)
    elif what == "prefill":
        prefill(4096, 14336, 8192)
        prefill(8192, 28672, 8192)
