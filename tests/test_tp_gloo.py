"""CPU, world_size 2 (gloo): host-side logic of the tensor-parallel feed-forward -- sharding, chunking, ragged
chunks, reduce-scatter / all-gather plumbing -- against the unsharded oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llama32_b200.tp import shard_range


def test_shard_ranges_cover_and_align():
    for inter in (14336, 28672, 688, 104, 1000):
        for world in (1, 2, 4, 8):
            if (inter, world) == (104, 8):   # 13 columns per rank round up to 16: the last rank would own nothing -> rejected
                import pytest
                with pytest.raises(ValueError):
                    [shard_range(inter, world, r) for r in range(world)]
                continue
            spans = [shard_range(inter, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == inter
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(lo % 8 == 0 for lo, _ in spans)
    assert all(shard_range(28672, 8, r) == (r * 3584, (r + 1) * 3584) for r in range(8))
    assert all((hi - lo) % 128 == 0 for lo, hi in (shard_range(14336, 8, r) for r in range(8)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tokens, hidden, inter, chunks, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import llama32_b200 as L
        from llama32_b200.tp import TensorParallelFFN
        from oracle import ffn_oracle as O
        torch.manual_seed(0)            # same weights and inputs on every rank
        ffn = L.FusedFeedforward(hidden, inter)
        x = torch.randn(2, tokens // 2, hidden) if tokens % 2 == 0 else torch.randn(tokens, hidden)
        ref = O.feedforward(x, ffn.swiglu.w_gate.detach(), ffn.swiglu.w_up.detach(), ffn.w_down.weight.detach())
        tp = TensorParallelFFN(ffn, chunks=chunks)
        y = tp(x)
        err = O.rel_l2(y, ref)
        # scattered hand-off: concatenating every rank's slices reproduces the full result
        pieces = tp.forward_scattered(x)
        rows_ok = all(p.shape[0] * world >= rows for _, rows, p in pieces)
        ret[rank] = (err, tuple(y.shape), rows_ok, tuple(tp.w_gate.shape), tuple(tp.w_down.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tokens,hidden,inter,chunks", [(16, 32, 256, 2), (13, 32, 104, 3), (64, 64, 688, 4)])
def test_tp_ffn_world2_matches_unsharded(tokens, hidden, inter, chunks):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), tokens, hidden, inter, chunks, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err, shape, rows_ok, wg_shape, wd_shape = ret[rank]
        assert err < 1e-5, (rank, err)
        assert rows_ok
        assert wg_shape[1] == hidden and wd_shape[0] == hidden and wg_shape[0] == wd_shape[1]
    assert ret[0][3][0] + ret[1][3][0] == inter
