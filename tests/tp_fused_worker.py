"""torchrun worker: fused tensor-parallel block on real GPUs (one process per GPU) against the CPU oracle.
Launched by tests/test_tp_fused_gpu.py::test_fused_tp_multi_process (and runnable by hand:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/tp_fused_worker.py).
Forward: small shapes (full oracle) and the shard widths of the 11B / 90B models at this world size (I/p = 1792 / 3584;
oracle on a row subset of every rank).  Forward + backward: autograd over the oracle, every gradient."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from llama32_b200.tp import FusedTensorParallelBlock, TpRankBuffers, shard_range  # noqa: E402
from oracle import ffn_oracle as O  # noqa: E402


def _check(got, ref, what, rtol=1e-2, mtol=2.0 ** -6):
    got, ref = got.float().cpu(), ref.float().cpu()
    r, m = O.rel_l2(got, ref), O.max_abs_over_max_ref(got, ref)
    assert r <= rtol and m <= mtol, f"{what}: rel-L2 {r:.3e}, max-abs/max|ref| {m:.3e}"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    bf = lambda t: t.to(dev, torch.bfloat16)
    # ---- forward, full oracle
    for tokens, hidden, inter in [(1024, 512, 2048), (4096, 1024, 4096), (777, 256, 1024)]:
        s = O.synthetic_ffn(tokens, hidden, inter, seed=11)   # same seed -> same tensors on every rank
        x, res, gamma, wg, wu, wd = (bf(s[k]) for k in ("x", "residual", "gamma", "w_gate", "w_up", "w_down"))
        bufs = TpRankBuffers.symmetric(tokens, hidden, torch.bfloat16, dev)
        blk = FusedTensorParallelBlock(gamma, 1e-5, wg, wu, wd, bufs)
        lo, hi, _ = blk.rows_of(tokens)
        ref = O.feedforward(O.add_rmsnorm(s["x"], s["gamma"], 1e-5, s["residual"]), s["w_gate"], s["w_up"], s["w_down"])
        for step in range(3):
            y = blk.forward(x[lo:hi], res[lo:hi], tokens)
            torch.cuda.synchronize()
            _check(y, ref[lo:hi], f"rank {rank} shape {(tokens, hidden, inter)} step {step}")
        dist.barrier()
    # ---- forward at the shard widths of the real models (n_act = 112 tiles at 1792), oracle on 48 rows of every rank
    for tokens, hidden, shard in [(2048, 4096, 1792), (2048, 2048, 3584)]:
        inter = shard * world
        s = O.synthetic_ffn(tokens, hidden, inter, seed=5)
        x, res, gamma, wg, wu, wd = (bf(s[k]) for k in ("x", "residual", "gamma", "w_gate", "w_up", "w_down"))
        bufs = TpRankBuffers.symmetric(tokens, hidden, torch.bfloat16, dev)
        blk = FusedTensorParallelBlock(gamma, 1e-5, wg, wu, wd, bufs)
        assert blk.w_gate.shape[0] == shard
        lo, hi, _ = blk.rows_of(tokens)
        rows = torch.arange(lo, hi)[torch.linspace(0, hi - lo - 1, 48).long()]
        ref = O.feedforward(O.add_rmsnorm(s["x"][rows], s["gamma"], 1e-5, s["residual"][rows]), s["w_gate"], s["w_up"], s["w_down"])
        for step in range(2):
            y = blk.forward(x[lo:hi], res[lo:hi], tokens)
            torch.cuda.synchronize()
            _check(y[rows - lo], ref, f"rank {rank} shard width {shard} step {step}")
        del bufs, blk
        dist.barrier()
    if rank == 0:
        print("tp fused ok", flush=True)
    # ---- forward + backward through autograd (apply), every gradient against autograd over the oracle
    for tokens, hidden, inter in [(1024, 512, 2048), (2048, 1024, 1792 * world if world <= 4 else 4096), (777, 256, 1024)]:
        s = O.synthetic_ffn(tokens, hidden, inter, seed=17)
        leaves = {k: s[k].clone().requires_grad_(True) for k in ("x", "residual", "gamma", "w_gate", "w_up", "w_down")}
        yref = O.feedforward(O.add_rmsnorm(leaves["x"], leaves["gamma"], 1e-5, leaves["residual"]), leaves["w_gate"],
                             leaves["w_up"], leaves["w_down"])
        yref.backward(s["dy"])
        bufs = TpRankBuffers.symmetric(tokens, hidden, torch.bfloat16, dev)
        blk = FusedTensorParallelBlock(bf(s["gamma"]), 1e-5, bf(s["w_gate"]), bf(s["w_up"]), bf(s["w_down"]), bufs)
        for t in (blk.gamma, blk.w_gate, blk.w_up, blk.w_down):
            t.requires_grad_(True)
        lo, hi, _ = blk.rows_of(tokens)
        slo, shi = shard_range(inter, world, rank)
        for step in range(2):
            for t in (blk.gamma, blk.w_gate, blk.w_up, blk.w_down):
                t.grad = None
            xl = bf(s["x"][lo:hi]).requires_grad_(True)
            rl = bf(s["residual"][lo:hi]).requires_grad_(True)
            y = blk.apply(xl, rl, tokens)
            y.backward(bf(s["dy"][lo:hi]))
            torch.cuda.synchronize()
            what = f"rank {rank} shape {(tokens, hidden, inter)} step {step}"
            _check(y, yref.detach()[lo:hi], what + " y")
            g = dict(rtol=1.5e-2, mtol=2.0 ** -5)
            _check(xl.grad, leaves["x"].grad[lo:hi], what + " dx", **g)
            _check(rl.grad, leaves["residual"].grad[lo:hi], what + " dresidual", **g)
            _check(blk.gamma.grad, leaves["gamma"].grad, what + " dgamma", **g)
            _check(blk.w_gate.grad, leaves["w_gate"].grad[slo:shi], what + " dw_gate", **g)
            _check(blk.w_up.grad, leaves["w_up"].grad[slo:shi], what + " dw_up", **g)
            _check(blk.w_down.grad, leaves["w_down"].grad[:, slo:shi], what + " dw_down", **g)
        del bufs, blk
        dist.barrier()
    if rank == 0:
        print("tp fused backward ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
