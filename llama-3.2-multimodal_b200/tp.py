"""Tensor-parallel SwiGLU feed-forward across the GPUs of one node (one process per GPU, torch.distributed).

The reference has no distributed code at all (grep for nccl / all_reduce / world_size finds
nothing); this is the sharding BASELINE.json's north_star asks for:
  * w_gate / w_up are column-parallel: rank r owns rows [r*I/p, (r+1)*I/p) of the [I, H] matrices
    (reference layout Tools/swiglu/FusedSwiglu.py:63-64), a contiguous slice -- no copy;
  * w_down is row-parallel: rank r owns columns [r*I/p, (r+1)*I/p) of the [H, I] matrix (reference
    Model/model.py:214), copied once to a contiguous [H, I/p] shard;
  * forward: x [M, H] replicated -> local fused gate/up+SiLU*mul GEMM -> local down GEMM gives a PARTIAL
    y_r [M, H]; the one real exchange step is reduce-scatter(y_r) over rows followed by all-gather
    (= all-reduce split in two so a sequence-parallel RMSNorm can sit in between).
The token dimension is processed in chunks: the NCCL reduce-scatter / all-gather of chunk c runs on NCCL's
stream while the tcgen05 GEMMs of chunk c+1 run on the compute stream, so the exchange over NVLink overlaps
the math.

Works on CPU tensors with the gloo backend too (tests/test_tp_gloo.py): the per-rank math then goes through the
same modules' fp32 expressions, and reduce-scatter is emulated with all_reduce + slice because gloo has no
reduce_scatter.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .modules import FusedFeedforward


def shard_range(inter: int, world: int, rank: int, granule: int = 128) -> tuple[int, int]:
    """Contiguous slice of the intermediate dimension owned by `rank`; multiples of `granule` (one act tile of
    the tcgen05 kernel) wherever possible so no rank gets a ragged tile."""
    units = inter // granule
    if units >= world:
        base, extra = divmod(units, world)
        lo = (rank * base + min(rank, extra)) * granule
        hi = lo + (base + (1 if rank < extra else 0)) * granule
        if rank == world - 1:
            hi = inter
        return lo, hi
    per = -(-inter // world)
    per = -(-per // 8) * 8
    lo = min(rank * per, inter)
    return lo, min(lo + per, inter)


def shard_ffn_weights(w_gate, w_up, w_down, world: int, rank: int):
    """(w_gate_r [I/p, H] view, w_up_r view, w_down_r [H, I/p] contiguous copy)."""
    lo, hi = shard_range(w_gate.shape[0], world, rank)
    return w_gate[lo:hi], w_up[lo:hi], w_down[:, lo:hi].contiguous()


class TensorParallelFFN(torch.nn.Module):
    """Drop-in for FusedFeedforward.forward on `group`: same input, same (replicated) output."""

    def __init__(self, ffn: FusedFeedforward, group=None, chunks: int = 4):
        super().__init__()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.chunks = chunks
        wg, wu, wd = shard_ffn_weights(ffn.swiglu.w_gate.detach(), ffn.swiglu.w_up.detach(), ffn.w_down.weight.detach(),
                                       self.world, self.rank)
        self.w_gate = torch.nn.Parameter(wg.contiguous(), requires_grad=False)
        self.w_up = torch.nn.Parameter(wu.contiguous(), requires_grad=False)
        self.w_down = torch.nn.Parameter(wd, requires_grad=False)
        self.hidden_size = self.w_gate.shape[1]

    # -- local math: sm_100a kernels for CUDA 16-bit tensors, the reference's fp32 expressions otherwise
    def _partial(self, x2):
        if ops.supported(x2):
            y, _, _ = ops.ffn_forward(x2, self.w_gate, self.w_up, self.w_down)
            return y
        act = torch.nn.functional.silu(torch.nn.functional.linear(x2, self.w_gate)) * torch.nn.functional.linear(x2, self.w_up)
        return torch.nn.functional.linear(act, self.w_down)

    def _reduce_scatter(self, part, out):
        if part.is_cuda:
            return dist.reduce_scatter_tensor(out, part, group=self.group, async_op=True)
        dist.all_reduce(part, group=self.group)          # gloo: no reduce_scatter
        rows = out.shape[0]
        out.copy_(part[self.rank * rows:(self.rank + 1) * rows])
        return None

    def forward_scattered(self, x):
        """Returns the list of (row_offset, rows, y_slice [rows/p, H]) per chunk: this rank's rows after the
        reduce-scatter -- the hand-off point for a sequence-parallel Add-RMSNorm."""
        x2 = x.reshape(-1, x.shape[-1])
        m = x2.shape[0]
        p = self.world
        step = -(-m // self.chunks)
        step = max(p, -(-step // p) * p)
        out, works = [], []
        for lo in range(0, m, step):
            hi = min(lo + step, m)
            rows = hi - lo
            part = self._partial(x2[lo:hi])
            if rows % p:                                   # ragged last chunk: pad rows so they split evenly
                pad = p - rows % p
                part = torch.cat([part, part.new_zeros(pad, part.shape[1])])
            piece = torch.empty(part.shape[0] // p, part.shape[1], dtype=part.dtype, device=part.device)
            works.append(self._reduce_scatter(part, piece))
            out.append((lo, rows, piece))
        for w in works:
            if w is not None:
                w.wait()
        return out

    def forward(self, x):
        x2 = x.reshape(-1, x.shape[-1])
        m, p = x2.shape[0], self.world
        y = torch.empty(m, self.hidden_size, dtype=x2.dtype, device=x2.device)
        works = []
        for lo, rows, piece in self.forward_scattered(x):
            padded = piece.shape[0] * p
            if padded == rows:
                works.append(dist.all_gather_into_tensor(y[lo:lo + rows], piece, group=self.group, async_op=True))
            else:
                full = torch.empty(padded, piece.shape[1], dtype=piece.dtype, device=piece.device)
                dist.all_gather_into_tensor(full, piece, group=self.group)
                y[lo:lo + rows].copy_(full[:rows])
        for w in works:
            w.wait()
        return y.view(x.shape)


# ======================================================================================================================
# Fused path: the collectives live INSIDE the tcgen05 GEMM kernels (peer memory over NVLink 5 / NVSwitch), no NCCL on
# the data path.  Sequence-parallel in, sequence-parallel out:
#
#   rank r holds rows [r*R, (r+1)*R) of x / residual (R = ceil(tokens / world))
#   1. Add-RMSNorm of the own rows, written into the own full-size `normed` buffer      (K1, local)
#      -> signal ready[r] = epoch on every rank
#   2. gate/up GEMM + SiLU*mul on the weight shard; the other ranks' normed rows are PULLED out of peer memory by the
#      spare warps of the GEMM CTAs while the tensor cores work on the rows already there (fused all-gather)
#   3. down GEMM on the shard; the epilogue stores every output row straight into the slot of the rank that owns the
#      row (fused reduce-scatter, peer stores)  -> signal rs_done[r] = epoch on every rank
#   4. the owner sums the `world` slots (fp32, rank order) once every peer has signalled -> y rows [r*R, (r+1)*R)
#
# Buffer reuse across steps is safe without extra barriers: a rank can only start step i+1's norm after its own step-i
# reduce, which waited for every peer's rs_done flag, which each peer raises after its step-i GEMMs (the pulls included).
# ======================================================================================================================
FLAG_READY, FLAG_RS_DONE, NUM_FLAGS = 0, 8, 64


class TpRankBuffers:
    """Buffers of one rank of the fused path plus the addresses of every rank's buffers (entry [rank] = own)."""

    def __init__(self, rank, world, max_tokens, hidden, normed, slots, flags, peer_normed, peer_slots_base, peer_flags):
        self.rank, self.world, self.max_tokens, self.hidden = rank, world, max_tokens, hidden
        self.slot_rows = slots.shape[1]
        self.normed, self.slots, self.flags = normed, slots, flags
        self.done = torch.zeros(8, dtype=torch.int32, device=normed.device)
        self.peer_normed = list(peer_normed)
        self.peer_flags = list(peer_flags)
        slot_bytes = self.slot_rows * hidden * normed.element_size()
        # where THIS rank's partial for owner o lands: slot [rank] of rank o's slots buffer
        self.peer_slots = [int(base) + rank * slot_bytes for base in peer_slots_base]

    @staticmethod
    def slot_rows_for(max_tokens, world):
        return -(-max_tokens // world)

    @classmethod
    def symmetric(cls, max_tokens, hidden, dtype, device, group=None):
        """Allocate the three buffers as symmetric memory and exchange the peer mappings (one process per GPU)."""
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        rows = cls.slot_rows_for(max_tokens, world)
        normed = symm.empty(max_tokens, hidden, dtype=dtype, device=device)
        slots = symm.empty(world, rows, hidden, dtype=dtype, device=device)
        flags = symm.empty(NUM_FLAGS, dtype=torch.int32, device=device)
        flags.zero_()
        hn, hs, hf = (symm.rendezvous(t, group) for t in (normed, slots, flags))
        torch.cuda.synchronize(device)
        dist.barrier(group)   # every rank's flags are zero before anybody can signal
        bufs = cls(rank, world, max_tokens, hidden, normed, slots, flags, hn.buffer_ptrs, hs.buffer_ptrs, hf.buffer_ptrs)
        bufs._handles = (hn, hs, hf)   # keep the mappings alive
        return bufs

    @classmethod
    def local_world(cls, world, max_tokens, hidden, dtype, device):
        """All ranks' buffers on ONE device, cross-wired with plain pointers: the single-GPU emulation used by the
        tests (the ranks' phases are then run one after another, never concurrently)."""
        rows = cls.slot_rows_for(max_tokens, world)
        normed = [torch.zeros(max_tokens, hidden, dtype=dtype, device=device) for _ in range(world)]
        slots = [torch.zeros(world, rows, hidden, dtype=dtype, device=device) for _ in range(world)]
        flags = [torch.zeros(NUM_FLAGS, dtype=torch.int32, device=device) for _ in range(world)]
        return [cls(r, world, max_tokens, hidden, normed[r], slots[r], flags[r], [t.data_ptr() for t in normed],
                    [t.data_ptr() for t in slots], [t.data_ptr() for t in flags]) for r in range(world)]


class FusedTensorParallelBlock:
    """norm2(x, residual) -> feed-forward of one rank, collectives fused into the GEMM kernels (see above)."""

    def __init__(self, gamma, eps, w_gate, w_up, w_down, bufs: TpRankBuffers, one_kernel: bool = False):
        """gamma [H]; w_gate / w_up / w_down are the FULL matrices ([I,H], [I,H], [H,I]); the shard is taken here.
        one_kernel: run gate/up and down as ONE persistent kernel (l32_tp_ffn_forward_fused) instead of two.  Measured
        at 8 GPUs: equal within 2-5 % either way -- the two-kernel path already hides both
        collectives -- so the simpler two-kernel path is the default."""
        self.bufs = bufs
        self.one_kernel = one_kernel
        self.eps = eps
        self.gamma = gamma
        wg, wu, wd = shard_ffn_weights(w_gate, w_up, w_down, bufs.world, bufs.rank)
        self.w_gate, self.w_up, self.w_down = wg.contiguous(), wu.contiguous(), wd
        self.epoch = 0
        self._act = None

    def rows_of(self, tokens, rank=None):
        rank = self.bufs.rank if rank is None else rank
        per = -(-tokens // self.bufs.world)
        lo = min(rank * per, tokens)
        return lo, min(lo + per, tokens), per

    # -- the four phases (run back to back by forward(); the single-GPU emulation interleaves them across ranks)
    def phase_norm(self, x_local, residual_local, tokens):
        b = self.bufs
        self.epoch += 1
        lo, hi, _ = self.rows_of(tokens)
        if hi > lo:
            ops.add_rmsnorm_forward(x_local, self.gamma, residual_local, self.eps, want_rms=False, out=b.normed[lo:hi])
        ops.tp_signal(b.peer_flags, FLAG_READY + b.rank, self.epoch, b.normed.device, zero8=b.done)

    def phase_gate_up(self, tokens):
        b = self.bufs
        _, _, per = self.rows_of(tokens)
        self._act = ops.tp_swiglu_forward_allgather(b.normed[:tokens], b.peer_normed, b.flags[FLAG_READY:FLAG_READY + 8],
                                                    b.done, self.epoch, b.rank, per, self.w_gate, self.w_up)

    def phase_down(self, tokens):
        b = self.bufs
        _, _, per = self.rows_of(tokens)
        ops.tp_linear_forward_reduce_scatter(self._act, self.w_down, b.peer_slots, b.rank, per)
        ops.tp_signal(b.peer_flags, FLAG_RS_DONE + b.rank, self.epoch, b.normed.device)

    def phase_ffn(self, tokens):
        """phase_gate_up + phase_down as one persistent kernel, then the rs_done signal."""
        b = self.bufs
        _, _, per = self.rows_of(tokens)
        ops.tp_ffn_forward_fused(b.normed[:tokens], b.peer_normed, b.flags[FLAG_READY:FLAG_READY + 8], b.done, self.epoch,
                                 b.rank, per, self.w_gate, self.w_up, self.w_down, b.peer_slots)
        ops.tp_signal(b.peer_flags, FLAG_RS_DONE + b.rank, self.epoch, b.normed.device)

    def phase_reduce(self, tokens, addend=None):
        b = self.bufs
        lo, hi, _ = self.rows_of(tokens)
        return ops.tp_reduce_partials(b.slots, b.flags[FLAG_RS_DONE:FLAG_RS_DONE + 8], self.epoch, b.rank, hi - lo, addend=addend)

    def forward(self, x_local, residual_local, tokens, addend=None):
        """x_local / residual_local: this rank's rows [rows_local, H]; returns this rank's rows of the FFN output."""
        if tokens > self.bufs.max_tokens:
            raise ValueError(f"tokens {tokens} exceeds the buffers' capacity {self.bufs.max_tokens}")
        self.phase_norm(x_local, residual_local, tokens)
        if self.one_kernel:
            self.phase_ffn(tokens)
        else:
            self.phase_gate_up(tokens)
            self.phase_down(tokens)
        return self.phase_reduce(tokens, addend)
