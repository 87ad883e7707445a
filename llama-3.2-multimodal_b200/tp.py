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
