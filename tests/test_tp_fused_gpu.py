"""Fused tensor-parallel path (collectives inside the tcgen05 GEMM kernels, peer memory): parity against the oracle.

* single-GPU emulation: all ranks' buffers live on one device and the ranks' phases run one after another (never
  concurrently -- B200_PROFILING.md forbids kernels that wait on each other as separate launches on one GPU);
* real multi-process run (torchrun, one process per GPU, symmetric memory over NVLink) when >= 2 GPUs are visible.
Tolerance: the forward gate of tests/test_gpu_parity.py (rel-L2 <= 1e-2, max-abs <= 2^-6 max|ref|); the partial sums are
rounded to bf16 once per rank before the owner adds them in fp32, like a bf16 NCCL reduce-scatter.
"""
import os
import subprocess
import sys

import pytest
import torch

from llama32_b200 import ops
from llama32_b200.tp import FusedTensorParallelBlock, TpRankBuffers
from oracle import ffn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _close(got, ref, what):
    got, ref = got.float().cpu(), ref.float().cpu()
    r, m = O.rel_l2(got, ref), O.max_abs_over_max_ref(got, ref)
    assert r <= 1e-2 and m <= 2.0 ** -6, f"{what}: rel-L2 {r:.3e}, max-abs/max|ref| {m:.3e}"


@pytest.mark.parametrize("one_kernel", [False, True])
@pytest.mark.parametrize("world,tokens,hidden,inter", [
    (1, 300, 256, 512), (2, 512, 512, 1024), (4, 1000, 256, 1536), (8, 2048, 1024, 2048), (3, 777, 384, 1152),
    (2, 4096, 512, 1792), (8, 8192, 256, 1024)])
def test_fused_tp_single_gpu_emulation(world, tokens, hidden, inter, one_kernel):
    s = O.synthetic_ffn(tokens, hidden, inter, seed=world * 100 + 7)
    bf = lambda t: t.to(DEV, torch.bfloat16)
    x, res, gamma, wg, wu, wd = (bf(s[k]) for k in ("x", "residual", "gamma", "w_gate", "w_up", "w_down"))
    bufs = TpRankBuffers.local_world(world, tokens, hidden, torch.bfloat16, DEV)
    blocks = [FusedTensorParallelBlock(gamma, 1e-5, wg, wu, wd, b) for b in bufs]
    ref = O.feedforward(O.add_rmsnorm(s["x"], s["gamma"], 1e-5, s["residual"]), s["w_gate"], s["w_up"], s["w_down"])
    for step in range(2):   # second step re-uses buffers and flags (epoch 2)
        for blk in blocks:
            lo, hi, _ = blk.rows_of(tokens)
            blk.phase_norm(x[lo:hi], res[lo:hi], tokens)
        if one_kernel:   # gate/up + down as ONE persistent kernel per rank
            for blk in blocks:
                blk.phase_ffn(tokens)
        else:
            for blk in blocks:
                blk.phase_gate_up(tokens)
            for blk in blocks:
                blk.phase_down(tokens)
        ys = [blk.phase_reduce(tokens) for blk in blocks]
        torch.cuda.synchronize()
        y = torch.cat(ys)
        assert y.shape == (tokens, hidden)
        _close(y, ref, f"fused TP world={world} step={step}")
    # the fused all-gather must have reproduced the full normalised activations on every rank
    normed_ref = O.add_rmsnorm(s["x"], s["gamma"], 1e-5, s["residual"])
    for b in bufs:
        _close(b.normed[:tokens], normed_ref, f"gathered activations on rank {b.rank}")
    # block-level fusion hook: y + addend in the reduction
    addend = torch.randn(tokens, hidden, device=DEV).bfloat16()
    for blk in blocks:
        lo, hi, _ = blk.rows_of(tokens)
        ya = blk.phase_reduce(tokens, addend=addend[lo:hi])
        _close(ya, ref[lo:hi] + addend[lo:hi].float().cpu(), "reduce + addend")


def _grad_close(got, ref, what):
    got, ref = got.float().cpu(), ref.float().cpu()
    r, m = O.rel_l2(got, ref), O.max_abs_over_max_ref(got, ref)
    assert r <= 1.5e-2 and m <= 2.0 ** -5, f"{what}: rel-L2 {r:.3e}, max-abs/max|ref| {m:.3e}"


def _oracle_block_grads(s, eps=1e-5):
    """autograd over the oracle: y = feedforward(add_rmsnorm(x, gamma, eps, residual)), upstream gradient dy."""
    leaves = {k: s[k].clone().requires_grad_(True) for k in ("x", "residual", "gamma", "w_gate", "w_up", "w_down")}
    y = O.feedforward(O.add_rmsnorm(leaves["x"], leaves["gamma"], eps, leaves["residual"]), leaves["w_gate"], leaves["w_up"],
                      leaves["w_down"])
    y.backward(s["dy"])
    return y.detach(), {k: v.grad for k, v in leaves.items()}


@pytest.mark.parametrize("world,tokens,hidden,inter", [
    (1, 300, 256, 512), (2, 512, 512, 1024), (4, 1000, 256, 1536), (8, 2048, 512, 2048), (3, 777, 384, 1152)])
def test_fused_tp_backward_single_gpu_emulation(world, tokens, hidden, inter):
    """Tensor-parallel forward + BACKWARD (all-gather of dY pulled inside the d_act GEMM, reduce-scatter of dX pushed from
    the two-phase dX GEMM, local weight gradients, RMSNorm backward on the own rows) against autograd over the oracle."""
    from llama32_b200.tp import shard_range
    s = O.synthetic_ffn(tokens, hidden, inter, seed=world * 10 + 3)
    bf = lambda t: t.to(DEV, torch.bfloat16)
    x, res, gamma, wg, wu, wd, dy = (bf(s[k]) for k in ("x", "residual", "gamma", "w_gate", "w_up", "w_down", "dy"))
    bufs = TpRankBuffers.local_world(world, tokens, hidden, torch.bfloat16, DEV)
    blocks = [FusedTensorParallelBlock(gamma, 1e-5, wg, wu, wd, b) for b in bufs]
    yref, gref = _oracle_block_grads(s)
    for step in range(2):
        saved = []
        for blk in blocks:
            from llama32_b200.tp import TpSaved
            sv = TpSaved()
            lo, hi, _ = blk.rows_of(tokens)
            blk.phase_norm(x[lo:hi], res[lo:hi], tokens, sv)
            saved.append(sv)
        for blk, sv in zip(blocks, saved):
            blk.phase_gate_up(tokens, sv)
        for blk in blocks:
            blk.phase_down(tokens)
        y = torch.cat([blk.phase_reduce(tokens) for blk in blocks])
        _close(y, yref, f"TP train forward world={world} step={step}")
        for blk, sv in zip(blocks, saved):
            lo, hi, _ = blk.rows_of(tokens)
            blk.bwd_phase_publish(dy[lo:hi], sv)
        for blk, sv in zip(blocks, saved):
            blk.bwd_phase_dact(sv)
        for blk, sv in zip(blocks, saved):
            blk.bwd_phase_dx(sv)
        wgr = [blk.bwd_phase_wgrads(sv) for blk, sv in zip(blocks, saved)]
        outs = [blk.bwd_phase_reduce_norm(sv) for blk, sv in zip(blocks, saved)]
        torch.cuda.synchronize()
        dx = torch.cat([o[0] for o in outs])
        dgamma = sum(o[1].float() for o in outs)
        _grad_close(dx, gref["x"], f"dx world={world} step={step}")
        _grad_close(dx, gref["residual"], "d_residual")
        _grad_close(dgamma, gref["gamma"], "dgamma")
        for r, (dwg, dwu, dwd) in enumerate(wgr):
            lo, hi = shard_range(inter, world, r)
            _grad_close(dwg, gref["w_gate"][lo:hi], f"dw_gate shard {r}")
            _grad_close(dwu, gref["w_up"][lo:hi], f"dw_up shard {r}")
            _grad_close(dwd, gref["w_down"][:, lo:hi], f"dw_down shard {r}")


def test_two_blocks_share_one_buffer_set():
    """Two layers on ONE TpRankBuffers: the step counter lives in the buffers, so the second block's waits
    cannot be satisfied by the first block's flags and its output is its own, not a stale one."""
    world, tokens, hidden, inter = 2, 512, 256, 1024
    sa, sb = O.synthetic_ffn(tokens, hidden, inter, seed=21), O.synthetic_ffn(tokens, hidden, inter, seed=22)
    bf = lambda t: t.to(DEV, torch.bfloat16)
    bufs = TpRankBuffers.local_world(world, tokens, hidden, torch.bfloat16, DEV)
    layers = []
    for s in (sa, sb):
        layers.append([FusedTensorParallelBlock(bf(s["gamma"]), 1e-5, bf(s["w_gate"]), bf(s["w_up"]), bf(s["w_down"]), b)
                       for b in bufs])
    for rep in range(2):
        for s, blocks in zip((sa, sb), layers):
            x, res = bf(s["x"]), bf(s["residual"])
            for blk in blocks:
                lo, hi, _ = blk.rows_of(tokens)
                blk.phase_norm(x[lo:hi], res[lo:hi], tokens)
            for blk in blocks:
                blk.phase_gate_up(tokens)
            for blk in blocks:
                blk.phase_down(tokens)
            y = torch.cat([blk.phase_reduce(tokens) for blk in blocks])
            ref = O.feedforward(O.add_rmsnorm(s["x"], s["gamma"], 1e-5, s["residual"]), s["w_gate"], s["w_up"], s["w_down"])
            _close(y, ref, f"layer sharing buffers, rep {rep}")
    assert all(b.epoch == 4 for b in bufs) and len({blk.epoch for blocks in layers for blk in blocks}) == 2


def _pow2_upto_device_count():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    out, p = [], 2
    while p <= max(n, 2):
        out.append(p)
        p *= 2
    return out


@pytest.mark.parametrize("nproc", _pow2_upto_device_count() or [2])
def test_fused_tp_multi_process(nproc):
    """Real processes, one per GPU, symmetric memory over NVLink: every power-of-two world size the box offers.  Forward
    at small shapes plus an 11B-shard (I/p = 1792) and a 90B-shard (I/p = 3584) shape, and forward + backward."""
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(29533 + nproc), os.path.join(ROOT, "tests", "tp_fused_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "tp fused ok" in r.stdout, r.stdout[-4000:]
    assert "tp fused backward ok" in r.stdout, r.stdout[-4000:]
