"""Per-phase timing of the fused tensor-parallel block (torchrun, one process per GPU).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/tp_breakdown.py [11b|90b] [tokens]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from llama32_b200.tp import FusedTensorParallelBlock, TpRankBuffers  # noqa: E402


def nvlink_kib(index):
    """Sum of the NVLink data counters of one GPU (KiB transmitted, KiB received) from `nvidia-smi nvlink -gt d`; None if
    the tool / counters are unavailable.  Read before and after the timed loop: the difference is the NVLink traffic of the
    loop as the hardware counted it (evidence for the bytes the fused kernels move, next to their algorithmic figure)."""
    import re
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
    except (OSError, subprocess.TimeoutExpired):
        return None
    tx = [int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
    rx = [int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
    if not tx or not rx:
        return None
    return sum(tx), sum(rx)


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "11b"
    tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    hidden, inter = (4096, 14336) if wl == "11b" else (8192, 28672)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dt = torch.bfloat16
    torch.manual_seed(0)
    gen = torch.Generator(device=dev).manual_seed(1)
    gamma = (1 + 0.1 * torch.randn(hidden, device=dev, generator=gen)).to(dt)
    wg = ((torch.rand(inter, hidden, device=dev, generator=gen) * 2 - 1) / hidden ** 0.5).to(dt)
    wu = ((torch.rand(inter, hidden, device=dev, generator=gen) * 2 - 1) / hidden ** 0.5).to(dt)
    wd = ((torch.rand(hidden, inter, device=dev, generator=gen) * 2 - 1) / inter ** 0.5).to(dt)
    bufs = TpRankBuffers.symmetric(tokens, hidden, dt, dev)
    one = os.environ.get("TP_ONE_KERNEL", "0") != "0"
    blk = FusedTensorParallelBlock(gamma, 1e-5, wg, wu, wd, bufs, one_kernel=one)
    del wg, wu, wd
    lo, hi, _ = blk.rows_of(tokens)
    xs = [torch.randn(hi - lo, hidden, device=dev, generator=gen).to(dt) for _ in range(2)]
    rs = [torch.randn(hi - lo, hidden, device=dev, generator=gen).to(dt) for _ in range(2)]
    names = ["norm+signal", "gate/up+allgather", "down+reduce-scatter+signal", "reduce"]
    acc = [0.0] * 4
    iters, warm = 30, 5
    for i in range(warm):
        blk.forward(xs[i % 2], rs[i % 2], tokens)
    # back-to-back steps (no barrier in between, the CPU runs ahead): the regime bench.py reports.  Events between the
    # phases give the steady-state per-phase times on each rank's GPU timeline.
    dist.barrier()
    torch.cuda.synchronize()
    nv0 = nvlink_kib(local)
    dist.barrier()            # reading the counters takes a different time on every rank: start the loop together
    torch.cuda.synchronize()
    evs = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        blk.phase_norm(xs[i % 2], rs[i % 2], tokens)
        ev[1].record()
        if one:
            blk.phase_ffn(tokens)
            ev[2].record()
        else:
            blk.phase_gate_up(tokens)
            ev[2].record()
            blk.phase_down(tokens)
        ev[3].record()
        blk.phase_reduce(tokens)
        ev[4].record()
        evs.append(ev)
    e1.record()
    torch.cuda.synchronize()
    nv1 = nvlink_kib(local)
    if nv0 is not None and nv1 is not None:
        hidden_b = hidden * 2
        alg = (world - 1) * (tokens // world) * hidden_b          # pulled per step (all-gather) = pushed per step (reduce-scatter)
        print(f"[nvlink] rank {rank}: tx {(nv1[0] - nv0[0]) / iters / 1024:.1f} MiB/step rx {(nv1[1] - nv0[1]) / iters / 1024:.1f} MiB/step "
              f"(algorithmic: {alg / 2 ** 20:.1f} MiB pulled + {alg / 2 ** 20:.1f} MiB pushed per step per rank)", flush=True)
    total = e0.elapsed_time(e1) / iters
    for ev in evs:
        for k in range(4):
            acc[k] += ev[k].elapsed_time(ev[k + 1])
    t = torch.tensor(acc + [total], device=dev, dtype=torch.float64)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    if rank == 0:
        fl_gu = 4.0 * tokens * hidden * inter / world
        fl_dn = 2.0 * tokens * hidden * inter / world
        for r, g in enumerate(gathered):
            g = g.tolist()
            ph = [v / iters for v in g[:4]]
            print(f"{wl} p={world} tokens={tokens} {'ONE-KERNEL ffn (phase 2 = gate/up+down+signal, phase 3 = 0)' if one else 'two kernels'} rank {r}: " + "  ".join(f"{n} {v * 1e3:.0f} us" for n, v in zip(names, ph)) +
                  f" | gate/up {fl_gu / ph[1] / 1e9:.0f} TF/s down {fl_dn / ph[2] / 1e9:.0f} TF/s | sum {sum(ph):.3f} ms | back-to-back {g[4]:.3f} ms "
                  f"= {tokens / g[4] / 1e3:.2f} M tok/s", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
