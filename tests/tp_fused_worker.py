"""torchrun worker: fused tensor-parallel block on real GPUs (one process per GPU) against the CPU oracle and against
the NCCL baseline path.  Launched by tests/test_tp_fused_gpu.py::test_fused_tp_multi_process."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from llama32_b200.tp import FusedTensorParallelBlock, TpRankBuffers  # noqa: E402
from oracle import ffn_oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for tokens, hidden, inter in [(1024, 512, 2048), (4096, 1024, 4096), (777, 256, 1024)]:
        s = O.synthetic_ffn(tokens, hidden, inter, seed=11)   # same seed -> same tensors on every rank
        bf = lambda t: t.to(dev, torch.bfloat16)
        x, res, gamma, wg, wu, wd = (bf(s[k]) for k in ("x", "residual", "gamma", "w_gate", "w_up", "w_down"))
        bufs = TpRankBuffers.symmetric(tokens, hidden, torch.bfloat16, dev)
        blk = FusedTensorParallelBlock(gamma, 1e-5, wg, wu, wd, bufs)
        lo, hi, _ = blk.rows_of(tokens)
        ref = O.feedforward(O.add_rmsnorm(s["x"], s["gamma"], 1e-5, s["residual"]), s["w_gate"], s["w_up"], s["w_down"])
        for step in range(3):
            y = blk.forward(x[lo:hi], res[lo:hi], tokens)
            torch.cuda.synchronize()
            got = y.float().cpu()
            r = O.rel_l2(got, ref[lo:hi])
            m = O.max_abs_over_max_ref(got, ref[lo:hi])
            assert r <= 1e-2 and m <= 2.0 ** -6, f"rank {rank} shape {(tokens, hidden, inter)} step {step}: {r:.3e} {m:.3e}"
        dist.barrier()
    if rank == 0:
        print("tp fused ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
