"""NVLink pull / push bandwidth of a plain SM copy kernel between two ranks (torchrun --nproc-per-node 2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.distributed._symmetric_memory as symm  # noqa: E402

from llama32_b200._lib import check, lib  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nbytes = 64 << 20
buf = symm.empty(nbytes, dtype=torch.uint8, device=dev)
h = symm.rendezvous(buf, dist.group.WORLD)
loc = torch.empty(nbytes, dtype=torch.uint8, device=dev)
peer = h.buffer_ptrs[(rank + 1) % world]
st = torch.cuda.current_stream().cuda_stream


def run(dst, src, ctas, warps, unroll, iters=10, seg=0):
    for _ in range(2):
        check(lib().l32_tp_peer_copy(dst, src, nbytes, ctas, warps, unroll, seg, st), "copy")
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        check(lib().l32_tp_peer_copy(dst, src, nbytes, ctas, warps, unroll, seg, st), "copy")
    e1.record()
    torch.cuda.synchronize()
    return nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


for ctas, warps, unroll in [(148, 2, 8), (148, 4, 8), (148, 8, 8), (32, 16, 8)]:
    pull = run(loc.data_ptr(), peer, ctas, warps, unroll)
    push = run(peer, loc.data_ptr(), ctas, warps, unroll)
    if rank == 0:
        print(f"ctas {ctas} warps {warps} unroll {unroll}: pull {pull:.0f} GB/s  push {push:.0f} GB/s (both ranks active, each direction)", flush=True)
for seg in (128, 256, 512, 1024, 2048, 8192):
    pull = run(loc.data_ptr(), peer, 148, 4, 8, seg=seg)
    push = run(peer, loc.data_ptr(), 148, 4, 8, seg=seg)
    loc2 = run(loc.data_ptr(), buf.data_ptr(), 148, 4, 8, seg=seg)
    if rank == 0:
        print(f"contiguous run {seg} B at 8 KiB pitch (148 CTAs x 4 warps): pull {pull:.0f} GB/s  push {push:.0f} GB/s  local copy {loc2:.0f} GB/s", flush=True)
dist.destroy_process_group()
