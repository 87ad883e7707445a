"""CPU checks of the tensor-parallel host logic (no GPU, no process group): row / shard partitions."""
import pytest
import torch

from llama32_b200.tp import FusedTensorParallelBlock, TpRankBuffers, shard_ffn_weights, shard_range, shard_state_dict_for_rank


@pytest.mark.parametrize("inter,world", [(14336, 8), (28672, 8), (14336, 3), (688, 2), (1152, 3), (104, 4), (16, 2)])
def test_shard_range_is_a_partition(inter, world):
    spans = [shard_range(inter, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == inter
    for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
        assert lo <= hi == lo2
    assert all((hi - lo) % 8 == 0 for lo, hi in spans[:-1])          # TMA row pitch / 16-byte stores
    if inter // 128 >= world:
        assert all(lo % 128 == 0 for lo, _ in spans)                  # shards start on an act tile of the tcgen05 kernel


def test_shard_range_rejects_empty_shards():
    """a rank that would own no columns is an error at shard time, not a shape failure inside the kernels."""
    with pytest.raises(ValueError):
        [shard_range(8, 2, r) for r in range(2)]
    with pytest.raises(ValueError):
        [shard_range(40, 8, r) for r in range(8)]


def test_epoch_counter_lives_in_the_buffers():
    class _T:   # stand-in for a device tensor: TpRankBuffers only reads shape / element_size / device here
        def __init__(self, shape): self.shape = shape; self.device = "cpu"
        def element_size(self): return 2
    import llama32_b200.tp as tp
    b = tp.TpRankBuffers(0, 2, 64, 32, _T((64, 32)), _T((2, 32, 32)), _T((64,)), [0, 0], [0, 4096], [0, 0])
    assert [b.next_epoch() for _ in range(3)] == [1, 2, 3]
    assert b.peer_slots == [0, 4096]          # rank 0 writes slot 0 of every owner


def test_shard_ffn_weights_reassemble():
    torch.manual_seed(0)
    wg, wu, wd = torch.randn(384, 64), torch.randn(384, 64), torch.randn(64, 384)
    x = torch.randn(5, 64)
    full = torch.nn.functional.linear(torch.nn.functional.silu(x @ wg.t()) * (x @ wu.t()), wd)
    acc = torch.zeros_like(full)
    for r in range(3):
        g, u, d = shard_ffn_weights(wg, wu, wd, 3, r)
        assert d.is_contiguous()
        acc += torch.nn.functional.linear(torch.nn.functional.silu(x @ g.t()) * (x @ u.t()), d)
    assert torch.allclose(acc, full, atol=1e-4)


@pytest.mark.parametrize("tokens,world", [(8192, 8), (777, 3), (5, 8), (1000, 4), (1, 2)])
def test_rows_of_covers_every_row_once(tokens, world):
    class _B:   # the only fields rows_of reads
        pass
    covered = []
    for r in range(world):
        blk = FusedTensorParallelBlock.__new__(FusedTensorParallelBlock)
        blk.bufs = _B()
        blk.bufs.rank, blk.bufs.world = r, world
        lo, hi, per = blk.rows_of(tokens)
        assert 0 <= lo <= hi <= tokens and hi - lo <= per == TpRankBuffers.slot_rows_for(tokens, world)
        covered += list(range(lo, hi))
    assert covered == list(range(tokens))


def test_load_time_packing_matches_device_side_sharding():
    """shard_state_dict_for_rank (host, load time) == shard_ffn_weights (device side) for every rank; other keys untouched."""
    torch.manual_seed(1)
    state = {"language_model.model.trf_blocks.0.ff.swiglu.w_gate": torch.randn(512, 64),
             "language_model.model.trf_blocks.0.ff.swiglu.w_up": torch.randn(512, 64),
             "language_model.model.trf_blocks.0.ff.w_down.weight": torch.randn(64, 512),
             "language_model.model.trf_blocks.0.norm2.weight": torch.randn(64),
             "language_model.lm_head.weight": torch.randn(100, 64)}
    pre = "language_model.model.trf_blocks.0."
    for world in (2, 4):
        for rank in range(world):
            sh = shard_state_dict_for_rank(state, world, rank)
            g, u, d = shard_ffn_weights(state[pre + "ff.swiglu.w_gate"], state[pre + "ff.swiglu.w_up"], state[pre + "ff.w_down.weight"],
                                        world, rank)
            assert torch.equal(sh[pre + "ff.swiglu.w_gate"], g) and torch.equal(sh[pre + "ff.swiglu.w_up"], u)
            assert torch.equal(sh[pre + "ff.w_down.weight"], d) and sh[pre + "ff.w_down.weight"].is_contiguous()
            assert sh[pre + "norm2.weight"] is state[pre + "norm2.weight"] and sh["language_model.lm_head.weight"] is state["language_model.lm_head.weight"]
