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
    hi = min(lo + per, inter)
    if hi <= lo:
        raise ValueError(f"intermediate size {inter} is too small to shard over {world} ranks: rank {rank} would own "
                         "no columns (every rank needs at least 8)")
    return lo, hi


def shard_ffn_weights(w_gate, w_up, w_down, world: int, rank: int):
    """(w_gate_r [I/p, H] view, w_up_r view, w_down_r [H, I/p] contiguous copy)."""
    lo, hi = shard_range(w_gate.shape[0], world, rank)
    return w_gate[lo:hi], w_up[lo:hi], w_down[:, lo:hi].contiguous()


def shard_state_dict_for_rank(state, world: int, rank: int):
    """Load-time packing: the reference's loader builds a full `converted_state` and hands it to
    load_state_dict (Model/utils.py:149-166); under tensor parallelism a rank only needs its slice of every feed-forward
    weight.  This cuts `*.swiglu.w_gate` / `*.swiglu.w_up` (row slices) and `*.w_down.weight` (column slice, made contiguous)
    down to rank `rank`'s shard while the tensors are still on the host (or memory-mapped safetensors slices), so the full
    matrices never reach the device and FusedTensorParallelBlock(..., presharded=True) takes them as they are.
    Every other entry is passed through untouched.  Returns a new dict."""
    out = {}
    for name, t in state.items():
        if name.endswith("swiglu.w_gate") or name.endswith("swiglu.w_up"):
            lo, hi = shard_range(t.shape[0], world, rank)
            out[name] = t[lo:hi].contiguous()
        elif name.endswith("w_down.weight") or name.endswith("w_down.linear.weight"):
            lo, hi = shard_range(t.shape[1], world, rank)
            out[name] = t[:, lo:hi].contiguous()
        else:
            out[name] = t
    return out


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
# Backward mirrors it: the own rows of dY are published in the same exchange buffer, the d_act GEMM (SiLU' epilogue) pulls
# the other ranks' dY rows, the two-phase dX GEMM pushes its partial rows to their owners, the weight gradients are local
# GEMMs on the shard, and the owner sums the dX partials and runs the RMSNorm backward on its own rows.
#
# Buffer reuse across steps is safe without extra barriers: a rank can only start step i+1's publish after its own step-i
# reduce, which waited for every peer's rs_done flag, which each peer raises after its step-i GEMMs (the pulls included).
# The step counter ("epoch") belongs to the BUFFERS, not to a block: any number of blocks (one per layer) may share one
# buffer set, every forward or backward pass through any of them is one more epoch.
# ======================================================================================================================
FLAG_READY, FLAG_RS_DONE, NUM_FLAGS = 0, 8, 64


class TpRankBuffers:
    """Buffers of one rank of the fused path plus the addresses of every rank's buffers (entry [rank] = own).
    `normed` is the exchange buffer (normalised activations in forward, dY in backward); `epoch` counts the passes."""

    def __init__(self, rank, world, max_tokens, hidden, normed, slots, flags, peer_normed, peer_slots_base, peer_flags,
                 group=None):
        self.rank, self.world, self.max_tokens, self.hidden = rank, world, max_tokens, hidden
        self.slot_rows = slots.shape[1]
        self.normed, self.slots, self.flags = normed, slots, flags
        self.done = torch.zeros(8, dtype=torch.int32, device=normed.device)
        self.peer_normed = list(peer_normed)
        self.peer_flags = list(peer_flags)
        self.group = group
        self.epoch = 0
        slot_bytes = self.slot_rows * hidden * normed.element_size()
        # where THIS rank's partial for owner o lands: slot [rank] of rank o's slots buffer
        self.peer_slots = [int(base) + rank * slot_bytes for base in peer_slots_base]

    def next_epoch(self) -> int:
        """One more pass (forward or backward, of whichever block) through these buffers.  Every rank calls this the same
        number of times in the same order, so the counters agree without communication."""
        self.epoch += 1
        return self.epoch

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
        bufs = cls(rank, world, max_tokens, hidden, normed, slots, flags, hn.buffer_ptrs, hs.buffer_ptrs, hf.buffer_ptrs,
                   group=group)
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


class TpSaved:
    """What one training forward of a rank leaves for its backward."""
    __slots__ = ("tokens", "h", "rms", "x_full", "gate", "up", "dy_full", "d_gate", "d_up", "act", "has_residual")


class FusedTensorParallelBlock:
    """norm2(x, residual) -> feed-forward of one rank, collectives fused into the GEMM kernels (see above).
    Inference: forward().  Training: forward_train() / backward(), or `apply()` which wires both into autograd."""

    def __init__(self, gamma, eps, w_gate, w_up, w_down, bufs: TpRankBuffers, one_kernel: bool = False, presharded: bool = False):
        """gamma [H]; w_gate / w_up / w_down are the FULL matrices ([I,H], [I,H], [H,I]); the shard is taken here -- unless
        `presharded`: then they already are this rank's shard ([I/p,H], [I/p,H], [H,I/p], see shard_state_dict_for_rank).
        one_kernel: run gate/up and down as ONE persistent kernel (l32_tp_ffn_forward_fused) instead of two.  Measured
        at 8 GPUs: equal within 2-5 % either way -- the two-kernel path already hides both
        collectives -- so the simpler two-kernel path is the default."""
        self.bufs = bufs
        self.one_kernel = one_kernel
        self.eps = eps
        self.gamma = gamma
        if presharded:
            wg, wu, wd = w_gate, w_up, w_down.contiguous()
        else:
            wg, wu, wd = shard_ffn_weights(w_gate, w_up, w_down, bufs.world, bufs.rank)
        self.w_gate, self.w_up, self.w_down = wg.contiguous(), wu.contiguous(), wd
        self._epoch = 0
        self._act = None

    @property
    def epoch(self):
        """Epoch of the pass this block is currently in (the counter itself lives in the shared buffers)."""
        return self._epoch

    def rows_of(self, tokens, rank=None):
        rank = self.bufs.rank if rank is None else rank
        per = -(-tokens // self.bufs.world)
        lo = min(rank * per, tokens)
        return lo, min(lo + per, tokens), per

    def _check_tokens(self, tokens):
        if tokens > self.bufs.max_tokens:
            raise ValueError(f"tokens {tokens} exceeds the buffers' capacity {self.bufs.max_tokens}")

    # -- the four forward phases (run back to back by forward(); the single-GPU emulation interleaves them across ranks)
    def phase_norm(self, x_local, residual_local, tokens, saved: TpSaved | None = None):
        b = self.bufs
        self._epoch = b.next_epoch()
        lo, hi, _ = self.rows_of(tokens)
        if hi > lo:
            train = saved is not None
            _, rms, h = ops.add_rmsnorm_forward(x_local, self.gamma, residual_local, self.eps, want_h=train, want_rms=train,
                                                out=b.normed[lo:hi])
            if train:
                saved.h, saved.rms = (h if h is not None else x_local), rms
        if saved is not None:
            saved.tokens, saved.has_residual = tokens, residual_local is not None
        ops.tp_signal(b.peer_flags, FLAG_READY + b.rank, self._epoch, b.normed.device, zero8=b.done)

    def phase_gate_up(self, tokens, saved: TpSaved | None = None):
        b = self.bufs
        _, _, per = self.rows_of(tokens)
        ready = b.flags[FLAG_READY:FLAG_READY + 8]
        if saved is None:
            self._act = ops.tp_swiglu_forward_allgather(b.normed[:tokens], b.peer_normed, ready, b.done, self._epoch, b.rank,
                                                        per, self.w_gate, self.w_up)
            return
        # training: gather into a tensor of its own (the exchange buffer is recycled by the next pass) and keep the caches
        if b.world > 1:
            saved.x_full = torch.empty(tokens, b.hidden, dtype=b.normed.dtype, device=b.normed.device)
        else:
            saved.x_full = b.normed[:tokens].clone()
        x_arg = saved.x_full if b.world > 1 else b.normed[:tokens]
        self._act, saved.gate, saved.up = ops.tp_swiglu_forward_allgather(x_arg, b.peer_normed, ready, b.done, self._epoch,
                                                                          b.rank, per, self.w_gate, self.w_up, want_cache=True)

    def phase_down(self, tokens):
        b = self.bufs
        _, _, per = self.rows_of(tokens)
        ops.tp_linear_forward_reduce_scatter(self._act, self.w_down, b.peer_slots, b.rank, per)
        self._act = None
        ops.tp_signal(b.peer_flags, FLAG_RS_DONE + b.rank, self._epoch, b.normed.device)

    def phase_ffn(self, tokens):
        """phase_gate_up + phase_down as one persistent kernel, then the rs_done signal."""
        b = self.bufs
        _, _, per = self.rows_of(tokens)
        ops.tp_ffn_forward_fused(b.normed[:tokens], b.peer_normed, b.flags[FLAG_READY:FLAG_READY + 8], b.done, self._epoch,
                                 b.rank, per, self.w_gate, self.w_up, self.w_down, b.peer_slots)
        ops.tp_signal(b.peer_flags, FLAG_RS_DONE + b.rank, self._epoch, b.normed.device)

    def phase_reduce(self, tokens, addend=None):
        b = self.bufs
        lo, hi, _ = self.rows_of(tokens)
        return ops.tp_reduce_partials(b.slots, b.flags[FLAG_RS_DONE:FLAG_RS_DONE + 8], self._epoch, b.rank, hi - lo, addend=addend)

    def forward(self, x_local, residual_local, tokens, addend=None):
        """x_local / residual_local: this rank's rows [rows_local, H]; returns this rank's rows of the FFN output."""
        self._check_tokens(tokens)
        self.phase_norm(x_local, residual_local, tokens)
        if self.one_kernel:
            self.phase_ffn(tokens)
        else:
            self.phase_gate_up(tokens)
            self.phase_down(tokens)
        return self.phase_reduce(tokens, addend)

    # -- training: forward that keeps what the backward needs, and the backward phases
    def forward_train(self, x_local, residual_local, tokens):
        """Like forward(); also returns the TpSaved record for backward()."""
        self._check_tokens(tokens)
        saved = TpSaved()
        self.phase_norm(x_local, residual_local, tokens, saved)
        self.phase_gate_up(tokens, saved)
        self.phase_down(tokens)
        return self.phase_reduce(tokens), saved

    def bwd_phase_publish(self, dy_local, saved: TpSaved):
        """Own rows of dY into the exchange buffer + ready signal (a new epoch)."""
        b = self.bufs
        self._epoch = b.next_epoch()
        lo, hi, _ = self.rows_of(saved.tokens)
        if hi > lo:
            b.normed[lo:hi].copy_(dy_local.reshape(hi - lo, b.hidden))
        ops.tp_signal(b.peer_flags, FLAG_READY + b.rank, self._epoch, b.normed.device, zero8=b.done)

    def bwd_phase_dact(self, saved: TpSaved, want_dw_down=True):
        """d_act = dY w_down_shard with the all-gather of dY pulled in; SiLU' recomputed in the epilogue."""
        b = self.bufs
        tokens = saved.tokens
        _, _, per = self.rows_of(tokens)
        if b.world > 1:
            saved.dy_full = torch.empty(tokens, b.hidden, dtype=b.normed.dtype, device=b.normed.device)
            dy_arg = saved.dy_full
        else:
            dy_arg = b.normed[:tokens]
            saved.dy_full = dy_arg
        saved.d_gate, saved.d_up, saved.act = ops.tp_ffn_backward_dact_allgather(
            dy_arg, b.peer_normed, b.flags[FLAG_READY:FLAG_READY + 8], b.done, self._epoch, b.rank, per, self.w_down,
            saved.gate, saved.up, want_act=want_dw_down)

    def bwd_phase_dx(self, saved: TpSaved):
        """Partial dX = d_gate Wg_shard + d_up Wu_shard, rows pushed to their owners; then the rs_done signal."""
        b = self.bufs
        _, _, per = self.rows_of(saved.tokens)
        ops.tp_ffn_backward_dx_reduce_scatter(saved.d_gate, saved.d_up, self.w_gate, self.w_up, b.peer_slots, b.rank, per)
        ops.tp_signal(b.peer_flags, FLAG_RS_DONE + b.rank, self._epoch, b.normed.device)

    def bwd_phase_wgrads(self, saved: TpSaved, want_gate_up=True, want_down=True):
        """Weight gradients of the shard: local MN-major x MN-major GEMMs (they run while the peers' partial dX rows are
        still in flight).  Returns (dw_gate [I/p,H], dw_up [I/p,H], dw_down [H,I/p])."""
        dwg = dwu = dwd = None
        if want_gate_up:
            dwg = ops.gemm(saved.d_gate, saved.x_full, a_mn_major=True, b_mn_major=True)
            dwu = ops.gemm(saved.d_up, saved.x_full, a_mn_major=True, b_mn_major=True)
        if want_down:
            dwd = ops.gemm(saved.dy_full, saved.act, a_mn_major=True, b_mn_major=True)
        return dwg, dwu, dwd

    def bwd_phase_reduce_norm(self, saved: TpSaved, want_dgamma=True):
        """Sum of the dX partials of the own rows, then the RMSNorm backward on them.  Returns (dx_local, dgamma) where
        dgamma covers the OWN rows only (sum it over the ranks: `allreduce_dgamma`); d_residual == dx_local."""
        b = self.bufs
        lo, hi, _ = self.rows_of(saved.tokens)
        d_normed = ops.tp_reduce_partials(b.slots, b.flags[FLAG_RS_DONE:FLAG_RS_DONE + 8], self._epoch, b.rank, hi - lo)
        if hi <= lo:
            return d_normed, (torch.zeros_like(self.gamma) if want_dgamma else None)
        return ops.rmsnorm_backward(d_normed, saved.h, self.gamma, saved.rms, want_dweight=want_dgamma)

    def allreduce_dgamma(self, dgamma):
        """gamma is replicated, its gradient is the sum over the ranks' rows (fp32 on the wire; [H] elements)."""
        b = self.bufs
        if b.world > 1 and b.group is not None:
            g32 = dgamma.float()
            dist.all_reduce(g32, group=b.group)
            return g32.to(dgamma.dtype)
        return dgamma

    def backward(self, dy_local, saved: TpSaved, want_dw_gate_up=True, want_dw_down=True, want_dgamma=True):
        """Returns (dx_local, dgamma, dw_gate_shard, dw_up_shard, dw_down_shard)."""
        self.bwd_phase_publish(dy_local, saved)
        self.bwd_phase_dact(saved, want_dw_down)
        self.bwd_phase_dx(saved)
        dwg, dwu, dwd = self.bwd_phase_wgrads(saved, want_dw_gate_up, want_dw_down)
        dx, dgamma = self.bwd_phase_reduce_norm(saved, want_dgamma)
        if dgamma is not None:
            dgamma = self.allreduce_dgamma(dgamma)
        return dx, dgamma, dwg, dwu, dwd

    def apply(self, x_local, residual_local, tokens):
        """Autograd-aware call: y_local = block(x_local, residual_local).  Gradients flow to x_local, residual_local and to
        self.gamma / self.w_gate / self.w_up / self.w_down when they require grad (make them leaf tensors or Parameters)."""
        tensors = (x_local, residual_local, self.gamma, self.w_gate, self.w_up, self.w_down)
        if not (torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)):
            return self.forward(x_local, residual_local, tokens)
        return _TpBlockFunction.apply(self, tokens, x_local, residual_local, self.gamma, self.w_gate, self.w_up, self.w_down)


class _TpBlockFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, block, tokens, x_local, residual_local, gamma, w_gate, w_up, w_down):
        y, saved = block.forward_train(x_local.detach(), None if residual_local is None else residual_local.detach(), tokens)
        ctx.block, ctx.saved_rec = block, saved
        return y

    @staticmethod
    def backward(ctx, dy_local):
        blk, saved = ctx.block, ctx.saved_rec
        _, _, nx, nr, ng, nwg, nwu, nwd = ctx.needs_input_grad
        dx, dgamma, dwg, dwu, dwd = blk.backward(dy_local.contiguous(), saved, want_dw_gate_up=(nwg or nwu), want_dw_down=nwd,
                                                 want_dgamma=ng)
        ctx.saved_rec = None
        d_res = dx if (nr and saved.has_residual) else None
        return None, None, (dx if nx else None), d_res, dgamma, (dwg if nwg else None), (dwu if nwu else None), dwd
