"""GPU parity tests proper: the CUDA path (through the C-ABI) against
  (1) the reference's own fp32 outputs (tests/golden, produced by oracle/make_golden.py from the real reference),
  (2) the CPU oracle on seeded bf16-representable inputs, forward and every gradient,
  (3) at BASELINE.json's full sizes: row/column subsets against the oracle plus size-independent properties.
Tolerances (bf16/fp16 storage, fp32 accumulation; SURVEY.md section 8c / BASELINE.md section 6):
  forward  : rel-L2 <= 1e-2 and max-abs <= 2^-6 * max|ref|
  gradients: rel-L2 <= 1e-2 and max-abs <= 2^-5 * max|ref|
"""
import math

import pytest
import torch

import llama32_b200 as L
from conftest import load_golden
from llama32_b200 import ops
from oracle import ffn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
FWD = (1e-2, 2.0 ** -6)
BWD = (1e-2, 2.0 ** -5)


def close(got, ref, tol, what=""):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert torch.isfinite(got).all(), what
    r, m = O.rel_l2(got, ref), O.max_abs_over_max_ref(got, ref)
    assert r <= tol[0] and m <= tol[1], f"{what}: rel-L2 {r:.3e} (<= {tol[0]}), max-abs/max|ref| {m:.3e} (<= {tol[1]:.3e})"


def dev(t, dtype=torch.bfloat16):
    return t.to(DEV, dtype)


# ----------------------------------------------------------------------------------------------- goldens
@pytest.mark.parametrize("tag", ["small", "odd", "cfg1"])
def test_rmsnorm_vs_reference_outputs(tag):
    g = load_golden(f"rmsnorm_{tag}.npz")
    w = dev(g["weight"])
    wr = w.float().cpu()      # gamma is not bf16-representable in the fixture: compare on what the kernel saw
    for res_key, res in (("nores", None), ("res", g["residual"])):
        x = dev(g["x"]).requires_grad_(True)
        r = None if res is None else dev(res).requires_grad_(True)
        wp = w.clone().requires_grad_(True)
        y = L.RMSNormFunction.apply(x, wp, g["eps"], r)
        y.backward(dev(g["grad_out"]))
        yr, dxr, dwr, drr = O.add_rmsnorm_grads(g["x"], wr, g["eps"], res, g["grad_out"])
        close(y, yr, FWD, f"y {tag} {res_key}")
        close(x.grad, dxr, BWD, "dx")
        close(wp.grad, dwr, BWD, "dweight")
        if r is not None:
            close(r.grad, drr, BWD, "dresidual")
        # and directly against the reference's numbers (gamma rounding included in the tolerance)
        close(y, g[f"y_{res_key}"], FWD, "y vs reference fixture")
        close(x.grad, g[f"dx_{res_key}"], BWD, "dx vs reference fixture")


@pytest.mark.parametrize("tag", ["small", "cfg1"])
def test_ffn_vs_reference_outputs(tag):
    g = load_golden(f"ffn_{tag}.npz")
    hidden, inter = g["w_gate"].shape[1], g["w_gate"].shape[0]
    ff = L.FusedFeedforward(hidden, inter).to(DEV, torch.bfloat16)
    with torch.no_grad():
        ff.swiglu.w_gate.copy_(dev(g["w_gate"])); ff.swiglu.w_up.copy_(dev(g["w_up"])); ff.w_down.weight.copy_(dev(g["w_down"]))
    x = dev(g["x"]).requires_grad_(True)
    act = ff.swiglu(x)
    y = ff(x)
    close(act, g["act"], FWD, "act")
    close(y, g["y"], FWD, "y")
    y.backward(dev(g["grad_out"]))
    close(x.grad, g["dx"], BWD, "dx")
    if "dw_gate" in g:
        close(ff.swiglu.w_gate.grad, g["dw_gate"], BWD, "dw_gate")
        close(ff.swiglu.w_up.grad, g["dw_up"], BWD, "dw_up")
        close(ff.w_down.weight.grad, g["dw_down"], BWD, "dw_down")


def test_ffn_bias_vs_reference_outputs():
    g = load_golden("ffn_bias.npz")
    ff = L.FusedFeedforward(64, 104, bias=True).to(DEV, torch.bfloat16)
    with torch.no_grad():
        ff.swiglu.w_gate.copy_(dev(g["w_gate"])); ff.swiglu.w_up.copy_(dev(g["w_up"])); ff.w_down.weight.copy_(dev(g["w_down"]))
        ff.swiglu.b_gate.copy_(dev(g["b_gate"])); ff.swiglu.b_up.copy_(dev(g["b_up"])); ff.w_down.bias.copy_(dev(g["b_down"]))
    bg, bu, bd = (ff.swiglu.b_gate.float().cpu(), ff.swiglu.b_up.float().cpu(), ff.w_down.bias.float().cpu())
    x = dev(g["x"]).requires_grad_(True)
    y = ff(x)
    yr = O.feedforward(g["x"], g["w_gate"], g["w_up"], g["w_down"], bg, bu, bd)
    close(y, yr, FWD, "y (bias)")
    y.backward(dev(g["grad_out"]))
    close(x.grad, g["dx"], BWD, "dx (bias)")
    assert ff.swiglu.b_gate.grad is not None and ff.w_down.bias.grad is not None


def test_block_hot_path_vs_reference_outputs():
    g = load_golden("block_cfg1.npz")
    norm2 = L.LLAMARMSNorm(256, eps=g["eps"]).to(DEV, torch.bfloat16)
    ff = L.FusedFeedforward(256, 688).to(DEV, torch.bfloat16)
    with torch.no_grad():
        norm2.weight.copy_(dev(g["norm2_weight"]))
        ff.swiglu.w_gate.copy_(dev(g["w_gate"])); ff.swiglu.w_up.copy_(dev(g["w_up"])); ff.w_down.weight.copy_(dev(g["w_down"]))
    attn = dev(g["attn_out"])
    hidden_before = dev(g["hidden"])
    keep = hidden_before.clone()
    normed = norm2(attn, residual=hidden_before)
    out = attn + ff(normed)
    assert torch.equal(hidden_before, keep), "the shipped wrapper must not mutate the residual (SURVEY.md 0.5)"
    close(normed, g["normed"], FWD, "normed")
    close(out, g["block_out"], FWD, "block_out")


def test_block_tail_fused_vs_reference_outputs():
    """l32_block_tail_forward (norm2 -> ff -> '+ attn_out' in the down epilogue) against the reference's own block output
    (golden from the real TransformerBlock tail; 32 tokens -> the small-M kernels)."""
    g = load_golden("block_cfg1.npz")
    attn, hidden = dev(g["attn_out"]), dev(g["hidden"])
    norm = L.LLAMARMSNorm(256, eps=g["eps"]).to(DEV, torch.bfloat16)
    ff = L.FusedFeedforward(256, 688).to(DEV, torch.bfloat16)
    with torch.no_grad():
        norm.weight.copy_(dev(g["norm2_weight"]))
        ff.swiglu.w_gate.copy_(dev(g["w_gate"])); ff.swiglu.w_up.copy_(dev(g["w_up"])); ff.w_down.weight.copy_(dev(g["w_down"]))
        out = L.block_tail(norm, ff, attn, hidden)
        composed = attn + ff(norm(attn, residual=hidden))
    close(out, g["block_out"], FWD, "fused block tail vs reference block output")
    close(out, composed, (4e-3, 2.0 ** -7), "fused vs composed modules")
    xg = attn.clone().requires_grad_(True)     # training: falls back to composing the modules, differentiable
    L.block_tail(norm, ff, xg, hidden).sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad.float()).all()


@pytest.mark.parametrize("tokens,hidden,inter", [(300, 256, 688), (1111, 512, 1536), (64, 1024, 2048)])
def test_block_tail_fused_vs_oracle(tokens, hidden, inter):
    """Same fusion through the tiled tcgen05 kernels (> 128 tokens) and the small-M kernels, against the oracle's
    restatement of Model/model.py:270-273."""
    s = O.synthetic_ffn(tokens, hidden, inter, seed=tokens)
    attn, res, gamma, wg, wu, wd = (dev(s[k]) for k in ("x", "residual", "gamma", "w_gate", "w_up", "w_down"))
    out = ops.block_tail_forward(attn, res, gamma, 1e-5, wg, wu, wd)
    _, _, ref = O.block_hot_path(s["x"], s["residual"], s["gamma"], 1e-5, s["w_gate"], s["w_up"], s["w_down"])
    close(out, ref, FWD, "block tail")
    out2 = ops.block_tail_forward(attn, None, gamma, 1e-5, wg, wu, wd)     # norm without a residual (final_norm style)
    _, _, ref2 = O.block_hot_path(s["x"], torch.zeros_like(s["x"]), s["gamma"], 1e-5, s["w_gate"], s["w_up"], s["w_down"])
    close(out2, ref2, FWD, "block tail, no residual")


def test_lora_vs_reference_outputs():
    g = load_golden("lora_small.npz")
    lin = L.Linear_LORA(176, 64, rank=16, alpha=32.0, dropout=0.0).to(DEV, torch.bfloat16)
    with torch.no_grad():
        lin.linear.weight.copy_(dev(g["w"])); lin.lora_a.weight.copy_(dev(g["lora_a"])); lin.lora_b.weight.copy_(dev(g["lora_b"]))
    close(lin(dev(g["x"])), g["y"], (2e-2, 2.0 ** -5), "lora y")   # rank-16 side path runs in bf16 torch


# ----------------------------------------------------------------------------------------------- oracle, seeded
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,c", [(1, 256), (3, 8), (17, 4096), (127, 1000), (5, 8192), (2, 16384), (9, 250), (4, 20000)])
def test_add_rmsnorm_shapes(dtype, rows, c):
    torch.manual_seed(rows * 131 + c)
    rnd = O.bf16_representable if dtype == torch.bfloat16 else O.fp16_representable
    x, r, g = rnd(torch.randn(rows, c)), rnd(torch.randn(rows, c)), rnd(torch.randn(rows, c))
    w = rnd(1 + 0.1 * torch.randn(c))
    for res in (None, r):
        y, rms, h = ops.add_rmsnorm_forward(dev(x, dtype), dev(w, dtype), None if res is None else dev(res, dtype), 1e-5,
                                            want_h=True)
        close(y, O.add_rmsnorm(x, w, 1e-5, res), FWD, "y")
        close(rms, O.rms_of(x, 1e-5, res), (1e-5, 1e-5), "rms")
        hh = x if res is None else x + res
        if res is not None:
            close(h, hh, FWD, "h")
        hq = hh.to(dtype).float()
        dx, dw = ops.rmsnorm_backward(dev(g, dtype), dev(hq, dtype), dev(w, dtype), rms)
        cdx, cdw = O.add_rmsnorm_grads_closed_form(hq, w, 1e-5, g)
        close(dx, cdx, BWD, "dx")
        close(dw, cdw, BWD, "dw")


def test_rmsnorm_edge_cases():
    w = torch.ones(64, device=DEV, dtype=torch.bfloat16)
    y, rms, _ = ops.add_rmsnorm_forward(torch.zeros(0, 64, device=DEV, dtype=torch.bfloat16), w, None, 1e-5)
    assert y.shape == (0, 64) and rms.shape == (0,)
    y, rms, _ = ops.add_rmsnorm_forward(torch.zeros(3, 64, device=DEV, dtype=torch.bfloat16), w, None, 1e-5)
    assert torch.isfinite(y).all() and (y == 0).all()
    assert torch.allclose(rms, torch.full((3,), math.sqrt(1e-5), device=DEV))
    # 3-D input, non-contiguous input, fp32 weight (default-constructed module) all accepted
    x = torch.randn(2, 5, 128, device=DEV).to(torch.bfloat16)
    n = L.LLAMARMSNorm(128).to(DEV)            # fp32 weight
    close(n(x), O.add_rmsnorm(x.float().cpu(), torch.ones(128), 1e-6), FWD, "3-D")
    xt = torch.randn(128, 6, device=DEV).to(torch.bfloat16).t()
    close(n(xt), O.add_rmsnorm(xt.float().cpu(), torch.ones(128), 1e-6), FWD, "non-contiguous")


def test_raw_rmsnorm_extension_abi():
    """rmsnorm.forward/backward keep the reference signature, including the in-place residual update."""
    import rmsnorm
    x = torch.randn(6, 256, device=DEV).to(torch.float16)
    r = torch.randn(6, 256, device=DEV).to(torch.float16)
    w = torch.randn(256, device=DEV).to(torch.float16)
    r0 = r.clone()
    out, rms = rmsnorm.forward(x, w, r, 1e-5)
    assert out.dtype == torch.float16 and rms.dtype == torch.float32 and rms.shape == (6,)
    close(r, x.float() + r0.float(), FWD, "residual := x + residual")
    close(out, O.add_rmsnorm(x.float().cpu(), w.float().cpu(), 1e-5, r0.float().cpu()), FWD, "out")
    dx, dw = rmsnorm.backward(torch.ones_like(x), r, w, rms)
    assert dx.shape == x.shape and dw.shape == (256,) and dw.dtype == torch.float16


SHAPES = [(1, 64, 104), (3, 256, 688), (100, 256, 688), (128, 128, 256), (129, 512, 264), (300, 1024, 2048), (520, 328, 776)]


@pytest.mark.parametrize("tokens,hidden,inter", SHAPES)
def test_ffn_forward_backward_vs_oracle(tokens, hidden, inter):
    s = O.synthetic_ffn(tokens, hidden, inter, seed=tokens + hidden)
    x, wg, wu, wd, dy = (dev(s[k]) for k in ("x", "w_gate", "w_up", "w_down", "dy"))
    y, gate, up = ops.ffn_forward(x, wg, wu, wd, want_cache=True)
    gr, ur = O.gate_up(s["x"], s["w_gate"], s["w_up"])
    close(gate, gr, FWD, "gate cache")
    close(up, ur, FWD, "up cache")
    ref = O.feedforward_grads(s["x"], s["w_gate"], s["w_up"], s["w_down"], s["dy"])
    close(y, ref["y"], FWD, "y")
    y2, _, _ = ops.ffn_forward(x, wg, wu, wd, want_cache=False)
    assert torch.equal(y, y2), "cache flag must not change the result"
    dx, dwg, dwu, dwd, _, _ = ops.ffn_backward(dy, x, wg, wu, wd, gate, up)
    close(dx, ref["dx"], BWD, "dx")
    close(dwg, ref["dw_gate"], BWD, "dw_gate")
    close(dwu, ref["dw_up"], BWD, "dw_up")
    close(dwd, ref["dw_down"], BWD, "dw_down")


@pytest.mark.parametrize("tokens,hidden,inter", [(5, 64, 104), (260, 256, 688)])
def test_swiglu_extension_abi_vs_oracle(tokens, hidden, inter):
    """swiglu_fused.forward / backward / forward_down with the reference's signatures (3-D x, None biases)."""
    import swiglu_fused
    s = O.synthetic_ffn(tokens, hidden, inter, seed=11)
    x = dev(s["x"]).view(1, tokens, hidden)
    wg, wu, wd = dev(s["w_gate"]), dev(s["w_up"]), dev(s["w_down"])
    out, gc, uc = swiglu_fused.forward(x, wg, wu, None, None)
    assert out.shape == (1, tokens, inter) and gc.shape == out.shape and uc.shape == out.shape
    ga = O.bf16_representable(torch.randn(1, tokens, inter))
    ref = O.swiglu_grads(s["x"].view(1, tokens, hidden), s["w_gate"], s["w_up"], ga)
    close(out, ref["act"], FWD, "act")
    gx, gwg, gwu = swiglu_fused.backward(dev(ga), x, wg, wu, gc, uc)
    close(gx, ref["dx"], BWD, "grad_x")
    close(gwg, ref["dw_gate"], BWD, "grad_w_gate")
    close(gwu, ref["dw_up"], BWD, "grad_w_up")
    yd = swiglu_fused.forward_down(x, wg, wu, wd)
    close(yd, O.feedforward(s["x"].view(1, tokens, hidden), s["w_gate"], s["w_up"], s["w_down"]), FWD, "forward_down")


def test_fp16_ffn():
    s = O.synthetic_ffn(200, 256, 512, seed=5)
    h = {k: O.fp16_representable(v) for k, v in s.items()}
    y, _, _ = ops.ffn_forward(*(dev(h[k], torch.float16) for k in ("x", "w_gate", "w_up", "w_down")))
    close(y, O.feedforward(h["x"], h["w_gate"], h["w_up"], h["w_down"]), (2e-3, 2.0 ** -9), "fp16 y")


def test_ffn_edge_cases():
    wg = torch.randn(104, 64, device=DEV).to(torch.bfloat16)
    wu, wd = torch.randn_like(wg), torch.randn(64, 104, device=DEV).to(torch.bfloat16)
    y, _, _ = ops.ffn_forward(torch.zeros(0, 64, device=DEV, dtype=torch.bfloat16), wg, wu, wd)
    assert y.shape == (0, 64)
    y, _, _ = ops.ffn_forward(torch.zeros(7, 64, device=DEV, dtype=torch.bfloat16), wg, wu, wd)
    assert (y == 0).all()
    with pytest.raises(RuntimeError):   # hidden not a multiple of 8: loud error, no silent fallback
        ops.swiglu_forward(torch.zeros(4, 60, device=DEV, dtype=torch.bfloat16),
                           torch.zeros(104, 60, device=DEV, dtype=torch.bfloat16),
                           torch.zeros(104, 60, device=DEV, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):   # fp32 tensors never reach the kernels
        ops.swiglu_forward(torch.zeros(4, 64, device=DEV), wg.float(), wu.float())


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_operand_layouts(cta_group, a_mn, b_mn):
    torch.manual_seed(0)
    m, n, k = 392, 520, 328
    a = O.bf16_representable(torch.randn(m, k))
    b = O.bf16_representable(torch.randn(n, k))
    a1 = O.bf16_representable(torch.randn(m, 200))
    b1 = O.bf16_representable(torch.randn(n, 200))
    lay = lambda t, mn: dev(t.t().contiguous() if mn else t)
    d = ops.gemm(lay(a, a_mn), lay(b, b_mn), a_mn_major=a_mn, b_mn_major=b_mn, cta_group=cta_group)
    close(d, a @ b.t(), (5e-3, 2.0 ** -7), "single phase")
    d2 = ops.gemm(lay(a, a_mn), lay(b, b_mn), a_mn_major=a_mn, b_mn_major=b_mn, a1=lay(a1, a_mn), b1=lay(b1, b_mn),
                  cta_group=cta_group)
    close(d2, a @ b.t() + a1 @ b1.t(), (5e-3, 2.0 ** -7), "two phase")


def test_module_autograd_end_to_end_and_lora():
    s = O.synthetic_ffn(384, 256, 688, seed=21)
    norm = L.LLAMARMSNorm(256, eps=1e-5).to(DEV, torch.bfloat16)
    ff = L.FusedFeedforward(256, 688).to(DEV, torch.bfloat16)
    with torch.no_grad():
        norm.weight.copy_(dev(s["gamma"]))
        ff.swiglu.w_gate.copy_(dev(s["w_gate"])); ff.swiglu.w_up.copy_(dev(s["w_up"])); ff.w_down.weight.copy_(dev(s["w_down"]))
    # LoRA surgery on w_down (reference README.md:179-188): frozen base on the tcgen05 GEMM, dX-only through it
    lo = L.Linear_LORA(688, 256, rank=16, alpha=32.0, dropout=0.0).to(DEV, torch.bfloat16)
    la = O.bf16_representable(torch.randn(16, 688) / 688 ** 0.5)
    lb = O.bf16_representable(0.05 * torch.randn(256, 16))
    with torch.no_grad():
        lo.linear.weight.copy_(ff.w_down.weight); lo.lora_a.weight.copy_(dev(la)); lo.lora_b.weight.copy_(dev(lb))
    ff.w_down = lo
    x = dev(s["x"]).requires_grad_(True)
    r = dev(s["residual"]).requires_grad_(True)
    y = ff(norm(x, residual=r))
    y.backward(dev(s["dy"]))
    xs, rs = s["x"].clone().requires_grad_(True), s["residual"].clone().requires_grad_(True)
    gm, wg, wu = (s[k].clone().requires_grad_(True) for k in ("gamma", "w_gate", "w_up"))
    las, lbs = la.clone().requires_grad_(True), lb.clone().requires_grad_(True)
    act = O.swiglu(O.add_rmsnorm(xs, gm, 1e-5, rs), wg, wu)
    yr = O.linear_lora(act, s["w_down"], las, lbs, 32.0, 16)
    yr.backward(s["dy"])
    tol_f, tol_b = (1.5e-2, 2.0 ** -5), (2e-2, 2.0 ** -4)     # the rank-16 side path is plain bf16 torch
    close(y, yr, tol_f, "lora y")
    close(x.grad, xs.grad, tol_b, "dx"); close(r.grad, rs.grad, tol_b, "dresidual")
    close(norm.weight.grad, gm.grad, tol_b, "dgamma")
    close(ff.swiglu.w_gate.grad, wg.grad, tol_b, "dw_gate"); close(ff.swiglu.w_up.grad, wu.grad, tol_b, "dw_up")
    close(lo.lora_a.weight.grad, las.grad, tol_b, "dlora_a"); close(lo.lora_b.weight.grad, lbs.grad, tol_b, "dlora_b")
    assert lo.linear.weight.grad is None


@pytest.mark.parametrize("tokens,hidden,inter,rank", [(300, 256, 688, 16), (64, 512, 1024, 8), (1030, 384, 1152, 64)])
def test_ffn_lora_fused_vs_oracle(tokens, hidden, inter, rank):
    """The LoRA adapter as a second accumulation phase of the down GEMM (forward) and of the d_act GEMM (backward),
    against autograd over the reference's expressions (FusedSwiglu.py:18-20 + Linear_LORA.forward, model.py:120-121)."""
    s = O.synthetic_ffn(tokens, hidden, inter, seed=rank + tokens)
    alpha = 32.0
    la = O.bf16_representable(torch.randn(rank, inter) / inter ** 0.5)
    lb = O.bf16_representable(0.05 * torch.randn(hidden, rank))
    x, wg, wu, wd, dy = (dev(s[k]) for k in ("x", "w_gate", "w_up", "w_down", "dy"))
    scale = alpha / rank
    lbs = dev(lb) * scale
    y, t, gate, up = ops.ffn_lora_forward(x, wg, wu, wd, dev(la), lbs, want_cache=True)
    xs = s["x"].clone().requires_grad_(True)
    wgs, wus = s["w_gate"].clone().requires_grad_(True), s["w_up"].clone().requires_grad_(True)
    las, lbs_ref = la.clone().requires_grad_(True), lb.clone().requires_grad_(True)
    act = O.swiglu(xs, wgs, wus)
    yr = O.linear_lora(act, s["w_down"], las, lbs_ref, alpha, rank)
    yr.backward(s["dy"])
    close(y, yr, FWD, "lora-fused y")
    close(t, act.detach() @ la.t(), (1.5e-2, 2.0 ** -5), "t = act A^T")
    dx, dwg, dwu, dla, dlbs = ops.ffn_lora_backward(dy, x, wg, wu, wd, dev(la), lbs, t, gate, up)
    close(dx, xs.grad, BWD, "dx")
    close(dwg, wgs.grad, BWD, "dw_gate")
    close(dwu, wus.grad, BWD, "dw_up")
    close(dla, las.grad, (1.5e-2, 2.0 ** -5), "dlora_a")
    close(dlbs.float() * scale, lbs_ref.grad, (1.5e-2, 2.0 ** -5), "dlora_b")
    # module routing: eval / p = 0 -> fused Function; training with dropout -> the unfused path (still correct in expectation)
    ff = L.FusedFeedforward(hidden, inter).to(DEV, torch.bfloat16)
    lo = L.Linear_LORA(inter, hidden, rank=rank, alpha=alpha, dropout=0.05).to(DEV, torch.bfloat16)
    with torch.no_grad():
        ff.swiglu.w_gate.copy_(wg); ff.swiglu.w_up.copy_(wu)
        lo.linear.weight.copy_(wd); lo.lora_a.weight.copy_(dev(la)); lo.lora_b.weight.copy_(dev(lb))
    ff.w_down = lo
    ff.eval()
    with torch.no_grad():
        close(ff(x), yr, FWD, "module (eval) y")
    ff.train()
    yt = ff(x)
    assert yt.shape == y.shape and torch.isfinite(yt.float()).all()


# ----------------------------------------------------------------------------------------------- full size
@pytest.mark.parametrize("hidden,inter", [(4096, 14336), (8192, 28672)])
def test_full_size_prefill_subsets_and_properties(hidden, inter):
    """BASELINE configs 2 / 5 (4 x 2048 tokens): the oracle checks a row subset (the FFN is row-independent) and
    a column subset of the weight gradients; properties cover the rest."""
    tokens = 8192
    gen = torch.Generator(device=DEV).manual_seed(hidden)
    rn = lambda *sh: torch.randn(*sh, device=DEV, generator=gen)
    ru = lambda *sh: torch.rand(*sh, device=DEV, generator=gen) * 2 - 1
    x, res, dy = rn(tokens, hidden).bfloat16(), rn(tokens, hidden).bfloat16(), rn(tokens, hidden).bfloat16()
    wg, wu = (ru(inter, hidden) / hidden ** 0.5).bfloat16(), (ru(inter, hidden) / hidden ** 0.5).bfloat16()
    wd = (ru(hidden, inter) / inter ** 0.5).bfloat16()
    gamma = (1 + 0.1 * rn(hidden)).bfloat16()
    normed, rms, h = ops.add_rmsnorm_forward(x, gamma, res, 1e-5, want_h=True)
    y, gate, up = ops.ffn_forward(normed, wg, wu, wd, want_cache=True)
    rows = torch.randperm(tokens, generator=torch.Generator().manual_seed(1))[:96].sort().values
    c = lambda t: t.float().cpu()
    close(normed[rows], O.add_rmsnorm(c(x[rows]), c(gamma), 1e-5, c(res[rows])), FWD, "normed rows")
    nr = c(normed[rows])
    ref = O.feedforward_grads(nr, c(wg), c(wu), c(wd), c(dy[rows]))
    close(y[rows], ref["y"], FWD, "y rows")
    dx, dwg, dwu, dwd, d_gate, d_up = ops.ffn_backward(dy, normed, wg, wu, wd, gate, up)
    close(dx[rows], ref["dx"], BWD, "dx rows")
    # weight-gradient column subset: dw_gate[cols, :] = d_gate[:, cols]^T normed  (full reduction over 8192 tokens)
    cols = torch.arange(0, inter, inter // 24)[:24]
    dgc = O.swiglu_grads_closed_form  # noqa: F841  (closed form documented in the oracle)
    g_c, u_c = c(normed) @ c(wg[cols]).t(), c(normed) @ c(wu[cols]).t()
    dact_c = c(dy) @ c(wd[:, cols])
    sg = torch.sigmoid(g_c)
    dg_c = dact_c * u_c * (sg * (1 + g_c * (1 - sg)))
    du_c = dact_c * (g_c * sg)
    close(dwg[cols], dg_c.t() @ c(normed), BWD, "dw_gate rows")
    close(dwu[cols], du_c.t() @ c(normed), BWD, "dw_up rows")
    close(dwd[:, cols], c(dy).t() @ (torch.nn.functional.silu(g_c) * u_c), BWD, "dw_down cols")
    # properties: row independence (a permutation of tokens permutes the output) and determinism
    perm = torch.randperm(tokens, device=DEV, generator=gen)
    y_p, _, _ = ops.ffn_forward(normed[perm].contiguous(), wg, wu, wd)
    assert torch.equal(y_p, y[perm]), "FFN rows must be independent of their position / tile"
    y_again, _, _ = ops.ffn_forward(normed, wg, wu, wd)
    assert torch.equal(y_again, y), "bitwise run-to-run determinism"
    # weight gradients are additive over token chunks (checksum of checksums)
    half = tokens // 2
    _, a_g, _, a_d, _, _ = ops.ffn_backward(dy[:half], normed[:half], wg, wu, wd, gate[:half], up[:half], want_dx=False)
    _, b_g, _, b_d, _, _ = ops.ffn_backward(dy[half:], normed[half:], wg, wu, wd, gate[half:], up[half:], want_dx=False)
    close(a_g.float() + b_g.float(), dwg, (6e-3, 2.0 ** -6), "dw_gate additivity")
    close(a_d.float() + b_d.float(), dwd, (6e-3, 2.0 ** -6), "dw_down additivity")
    # RMSNorm backward at full size against the closed form on the row subset (+ dgamma via a 1024-row slice)
    dxn, dgam = ops.rmsnorm_backward(dy, h, gamma, rms)
    cdx, _ = O.add_rmsnorm_grads_closed_form(c(h[rows]), c(gamma), 1e-5, c(dy[rows]))
    close(dxn[rows], cdx, BWD, "rmsnorm dx rows")
    _, dgam_slice = ops.rmsnorm_backward(dy[:1024], h[:1024], gamma, rms[:1024])
    _, cdw = O.add_rmsnorm_grads_closed_form(c(h[:1024]), c(gamma), 1e-5, c(dy[:1024]))
    close(dgam_slice, cdw, BWD, "dgamma 1024 rows")
    assert torch.isfinite(dgam.float()).all()


def test_full_size_lora_step_11b():
    """BASELINE config 4: LoRA (rank 16, alpha 32) on a frozen w_down at the 11B shape, seq 2048, forward + backward with
    the adapter fused into the GEMMs.  Row subset against the oracle (autograd), adapter gradients against the closed
    form evaluated in fp32 on the device over all 2048 tokens."""
    tokens, hidden, inter, rank, alpha = 2048, 4096, 14336, 16, 32.0
    scale = alpha / rank
    gen = torch.Generator(device=DEV).manual_seed(4)
    rn = lambda *sh: torch.randn(*sh, device=DEV, generator=gen)
    ru = lambda *sh: torch.rand(*sh, device=DEV, generator=gen) * 2 - 1
    x, dy = rn(tokens, hidden).bfloat16(), rn(tokens, hidden).bfloat16()
    wg, wu = (ru(inter, hidden) / hidden ** 0.5).bfloat16(), (ru(inter, hidden) / hidden ** 0.5).bfloat16()
    wd = (ru(hidden, inter) / inter ** 0.5).bfloat16()
    la = (rn(rank, inter) / inter ** 0.5).bfloat16()
    lb = (0.02 * rn(hidden, rank)).bfloat16()
    lbs = (lb.float() * scale).bfloat16()
    y, t, gate, up = ops.ffn_lora_forward(x, wg, wu, wd, la, lbs, want_cache=True)
    dx, dwg, dwu, dla, dlbs = ops.ffn_lora_backward(dy, x, wg, wu, wd, la, lbs, t, gate, up)
    c = lambda v: v.float().cpu()
    rows = torch.randperm(tokens, generator=torch.Generator().manual_seed(2))[:48].sort().values
    xs = c(x[rows]).requires_grad_(True)
    yr = O.linear_lora(O.swiglu(xs, c(wg), c(wu)), c(wd), c(la), c(lbs) / scale, alpha, rank)
    yr.backward(c(dy[rows]))
    close(y[rows], yr, FWD, "lora y rows")
    close(dx[rows], xs.grad, BWD, "lora dx rows")
    f = lambda v: v.float()
    g_, u_ = f(x) @ f(wg).t(), f(x) @ f(wu).t()
    act = torch.nn.functional.silu(g_) * u_
    uu = f(dy) @ f(lbs)                                            # [T, rank] = dy (s B)
    close(t, act @ f(la).t(), (1.5e-2, 2.0 ** -5), "t = act A^T")
    close(dla, uu.t() @ act, (1.5e-2, 2.0 ** -5), "dlora_a")
    close(dlbs, f(dy).t() @ (act @ f(la).t()), (1.5e-2, 2.0 ** -5), "dlora_b (scaled matrix)")
    d_act = f(dy) @ f(wd) + uu @ f(la)
    sg = torch.sigmoid(g_)
    cols = torch.arange(0, inter, inter // 16)[:16].to(DEV)
    dg = d_act * u_ * (sg * (1 + g_ * (1 - sg)))
    close(dwg[cols], dg[:, cols].t() @ f(x), BWD, "dw_gate rows (through the fused adapter)")


@pytest.mark.parametrize("batch", [1, 2, 8, 33, 64])
def test_decode_shapes_11b(batch):
    """BASELINE config 3: KV-cached decode, x [B, 1, 4096]; whole-matrix check against the oracle."""
    hidden, inter = 4096, 14336
    gen = torch.Generator(device=DEV).manual_seed(batch)
    x = torch.randn(batch, 1, hidden, device=DEV, generator=gen).bfloat16()
    wg = ((torch.rand(inter, hidden, device=DEV, generator=gen) * 2 - 1) / 64).bfloat16()
    wu = ((torch.rand(inter, hidden, device=DEV, generator=gen) * 2 - 1) / 64).bfloat16()
    wd = ((torch.rand(hidden, inter, device=DEV, generator=gen) * 2 - 1) / 120).bfloat16()
    y, _, _ = ops.ffn_forward(x, wg, wu, wd)
    assert y.shape == (batch, 1, hidden)
    ref = torch.nn.functional.linear(
        torch.nn.functional.silu(x.float() @ wg.float().t()) * (x.float() @ wu.float().t()), wd.float())
    close(y, ref, FWD, "decode y")   # fp32 on the device: same expression as the oracle (FusedSwiglu.py:18-20)
    yo = O.feedforward(x[:1].float().cpu(), wg.float().cpu(), wu.float().cpu(), wd.float().cpu())
    close(y[:1], yo, FWD, "decode y[0] vs CPU oracle")


def test_decode_shape_90b():
    """KV-cached decode at the 90B shape (hidden 8192, hidden_dim 28672; 1.4 GB of weights streamed per step)."""
    hidden, inter, batch = 8192, 28672, 4
    gen = torch.Generator(device=DEV).manual_seed(9)
    x = torch.randn(batch, 1, hidden, device=DEV, generator=gen).bfloat16()
    wg = ((torch.rand(inter, hidden, device=DEV, generator=gen) * 2 - 1) / 90).bfloat16()
    wu = ((torch.rand(inter, hidden, device=DEV, generator=gen) * 2 - 1) / 90).bfloat16()
    wd = ((torch.rand(hidden, inter, device=DEV, generator=gen) * 2 - 1) / 170).bfloat16()
    y, _, _ = ops.ffn_forward(x, wg, wu, wd)
    xf = x.float().view(batch, hidden)
    ref = torch.nn.functional.linear(torch.nn.functional.silu(xf @ wg.float().t()) * (xf @ wu.float().t()), wd.float())
    close(y.view(batch, hidden), ref, FWD, "decode 90B y")


@pytest.mark.parametrize("tokens,hidden,inter,dtype,bias", [
    (1, 256, 688, torch.bfloat16, False),      # config-1 shape: ragged last row block (688 = 10.75 * 64)
    (3, 264, 696, torch.bfloat16, True),       # K tail (264 = 4.125 * 64), biases
    (16, 512, 1024, torch.float16, False),     # token count == UMMA N
    (17, 512, 1536, torch.bfloat16, True),     # first N = 32
    (65, 1024, 2048, torch.bfloat16, False),   # first N = 128
    (128, 768, 3072, torch.float16, True),     # largest decode batch
    (5, 8192, 3584, torch.bfloat16, False),    # 90B tensor-parallel shard (I / 8), K splits on both kernels
])
def test_decode_kernel_small_m(tokens, hidden, inter, dtype, bias):
    """K5 weight-streaming kernels (swap-AB tcgen05 + cluster split-K) against the fp32 expression of the
    reference's live path (FusedSwiglu.py:18-20 + model.py:217) on the same 16-bit inputs."""
    gen = torch.Generator(device=DEV).manual_seed(tokens * 7 + hidden)
    mk = lambda *s, scale=1.0: ((torch.rand(*s, device=DEV, generator=gen) * 2 - 1) * scale).to(dtype)
    x = torch.randn(tokens, hidden, device=DEV, generator=gen).to(dtype)
    wg, wu = mk(inter, hidden, scale=hidden ** -0.5), mk(inter, hidden, scale=hidden ** -0.5)
    wd = mk(hidden, inter, scale=inter ** -0.5)
    bg = mk(inter, scale=0.5) if bias else None
    bu = mk(inter, scale=0.5) if bias else None
    bd = mk(hidden, scale=0.5) if bias else None
    f = lambda t: None if t is None else t.float()
    act, _, _ = ops.swiglu_forward(x, wg, wu, bg, bu)
    act_ref = torch.nn.functional.silu(torch.nn.functional.linear(f(x), f(wg), f(bg))) * torch.nn.functional.linear(f(x), f(wu), f(bu))
    close(act, act_ref, FWD, "decode act")
    y, _, _ = ops.ffn_forward(x, wg, wu, wd, bg, bu, bd)
    y_ref = torch.nn.functional.linear(act_ref, f(wd), f(bd))
    close(y, y_ref, FWD, "decode y")
    # the down projection alone, on an input that is not an FFN intermediate
    a = torch.randn(tokens, inter, device=DEV, generator=gen).to(dtype)
    close(ops.linear_forward(a, wd, bd), torch.nn.functional.linear(f(a), f(wd), f(bd)), FWD, "decode linear")
    # small-M and tiled kernels must agree: the same rows through a > 128-token call take the tiled path
    big = torch.cat([x, torch.randn(200, hidden, device=DEV, generator=gen).to(dtype)])
    y_big, _, _ = ops.ffn_forward(big, wg, wu, wd, bg, bu, bd)
    close(y, y_big[:tokens], (4e-3, 2.0 ** -7), "small-M kernel vs tiled kernel")


def test_decode_is_deterministic_and_graph_capturable():
    """No atomics / no allocation inside the C-ABI: bitwise repeatable, and capturable in a CUDA graph
    (programmatic-dependent-launch edges included)."""
    hidden, inter, tokens = 1024, 4096, 8
    gen = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(tokens, hidden, device=DEV, generator=gen).bfloat16()
    wg = ((torch.rand(inter, hidden, device=DEV, generator=gen) * 2 - 1) / 32).bfloat16()
    wu = ((torch.rand(inter, hidden, device=DEV, generator=gen) * 2 - 1) / 32).bfloat16()
    wd = ((torch.rand(hidden, inter, device=DEV, generator=gen) * 2 - 1) / 64).bfloat16()
    gamma = torch.ones(hidden, device=DEV).bfloat16()
    y0 = ops.ffn_forward(ops.add_rmsnorm_forward(x, gamma, None, 1e-5)[0], wg, wu, wd)[0].clone()
    for _ in range(3):
        y1 = ops.ffn_forward(ops.add_rmsnorm_forward(x, gamma, None, 1e-5)[0], wg, wu, wd)[0]
        assert torch.equal(y0, y1)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            ops.ffn_forward(ops.add_rmsnorm_forward(x, gamma, None, 1e-5)[0], wg, wu, wd)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        yg = ops.ffn_forward(ops.add_rmsnorm_forward(x, gamma, None, 1e-5)[0], wg, wu, wd)[0]
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y0, yg)


def test_second_device_in_the_same_process():
    """Kernel attributes and SM counts are tracked per device: the hot path must work on cuda:1 after cuda:0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for d in ("cuda:0", "cuda:1", "cuda:0"):
        gen = torch.Generator(device=d).manual_seed(3)
        x = torch.randn(300, 512, device=d, generator=gen).bfloat16()
        wg = ((torch.rand(1024, 512, device=d, generator=gen) * 2 - 1) / 22).bfloat16()
        wu = ((torch.rand(1024, 512, device=d, generator=gen) * 2 - 1) / 22).bfloat16()
        wd = ((torch.rand(512, 1024, device=d, generator=gen) * 2 - 1) / 32).bfloat16()
        gamma = torch.ones(512, device=d).bfloat16()
        for rows in (300, 7):   # tiled kernels and small-M kernels
            xx = x[:rows]
            y = ops.ffn_forward(ops.add_rmsnorm_forward(xx, gamma, None, 1e-5)[0], wg, wu, wd)[0]
            n = O.add_rmsnorm(xx.float().cpu(), gamma.float().cpu(), 1e-5)
            close(y, O.feedforward(n, wg.float().cpu(), wu.float().cpu(), wd.float().cpu()), FWD, f"ffn on {d}")
        dy = torch.randn(300, 512, device=d, generator=gen).bfloat16()
        ops.rmsnorm_backward(dy, x, gamma, torch.ones(300, device=d))


def test_reference_cuda_rmsnorm_ab():
    """A/B against the one reference CUDA kernel that builds (oracle/_ref, fp16 only)."""
    import glob
    import importlib.util
    import os
    so = glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "rmsnorm_ref*.so"))
    if not so:
        pytest.skip("oracle/_ref not built")
    spec = importlib.util.spec_from_file_location("rmsnorm_ref", so[0])
    ref_ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_ext)
    torch.manual_seed(0)
    x = torch.randn(512, 4096, device=DEV).half()
    r = torch.randn(512, 4096, device=DEV).half()
    w = (1 + 0.1 * torch.randn(4096, device=DEV)).half()
    out_ref, rms_ref = ref_ext.forward(x, w, r.clone(), 1e-5)
    y, rms, _ = ops.add_rmsnorm_forward(x, w, r, 1e-5)
    oracle = O.add_rmsnorm(x.float().cpu(), w.float().cpu(), 1e-5, r.float().cpu())
    close(y, oracle, (2e-3, 2.0 ** -9), "ours vs oracle (fp16)")
    close(out_ref, oracle, (4e-3, 2.0 ** -8), "reference kernel vs oracle (fp16, it rounds the add)")
    close(y, out_ref, (4e-3, 2.0 ** -8), "ours vs reference kernel")
    close(rms, rms_ref, (1e-3, 1e-3), "rms")
    assert O.rel_l2(y.float().cpu(), oracle) <= O.rel_l2(out_ref.float().cpu(), oracle) * 1.05


# ----------------------------------------------------------------------------------------------- f1 / f2 rows of SURVEY 8(f)
@pytest.mark.parametrize("tokens,hidden,inter,with_res,chained", [
    (320, 256, 688, True, False), (320, 256, 688, True, True), (1000, 512, 1024, False, True), (64, 256, 688, True, True)])
def test_block_tail_autograd_and_chained_norm(tokens, hidden, inter, with_res, chained):
    """BlockTailFunction: out = attn_out + ff(norm2(attn_out, residual)) [+ the next block's norm1 in the same call],
    forward and every gradient against autograd over the reference's expressions (Model/model.py:267-273, :346)."""
    s = O.synthetic_ffn(tokens, hidden, inter, seed=tokens + inter)
    g = torch.Generator().manual_seed(5)
    gamma2 = O.bf16_representable(1 + 0.1 * torch.randn(hidden, generator=g))
    d_next = O.bf16_representable(torch.randn(tokens, hidden, generator=g))
    norm2 = L.LLAMARMSNorm(hidden, eps=1e-5).to(DEV, torch.bfloat16)
    nxt = L.LLAMARMSNorm(hidden, eps=1e-5).to(DEV, torch.bfloat16)
    ff = L.FusedFeedforward(hidden, inter).to(DEV, torch.bfloat16)
    with torch.no_grad():
        norm2.weight.copy_(dev(s["gamma"])); nxt.weight.copy_(dev(gamma2))
        ff.swiglu.w_gate.copy_(dev(s["w_gate"])); ff.swiglu.w_up.copy_(dev(s["w_up"])); ff.w_down.weight.copy_(dev(s["w_down"]))
    a = dev(s["x"]).requires_grad_(True)
    r = dev(s["residual"]).requires_grad_(True) if with_res else None
    out = L.block_tail(norm2, ff, a, r, next_norm=nxt if chained else None)
    loss_terms = [(out, dev(s["dy"]))]
    if chained:
        assert getattr(out, "_l32_prenormed", None) is not None and out._l32_prenormed[0] is nxt
        nn_ = nxt(out)                                   # consumed from the attachment, no second kernel
        assert nn_ is out._l32_prenormed[1]
        loss_terms.append((nn_, dev(d_next)))
    torch.autograd.backward([t for t, _ in loss_terms], [gr for _, gr in loss_terms])
    # oracle
    leaves = {k: s[k].clone().requires_grad_(True) for k in ("x", "residual", "gamma", "w_gate", "w_up", "w_down")}
    g2 = gamma2.clone().requires_grad_(True)
    normed = O.add_rmsnorm(leaves["x"], leaves["gamma"], 1e-5, leaves["residual"] if with_res else None)
    out_r = leaves["x"] + O.feedforward(normed, leaves["w_gate"], leaves["w_up"], leaves["w_down"])
    terms = [(out_r, s["dy"])]
    if chained:
        terms.append((O.add_rmsnorm(out_r, g2, 1e-5), d_next))
    torch.autograd.backward([t for t, _ in terms], [gr for _, gr in terms])
    close(out, out_r, FWD, "block out")
    if chained:
        close(nn_, terms[1][0], FWD, "chained next norm")
        close(nxt.weight.grad, g2.grad, BWD, "d next gamma")
    close(a.grad, leaves["x"].grad, BWD, "d attn_out")
    if with_res:
        close(r.grad, leaves["residual"].grad, BWD, "d residual")
    close(norm2.weight.grad, leaves["gamma"].grad, BWD, "d gamma")
    close(ff.swiglu.w_gate.grad, leaves["w_gate"].grad, BWD, "dw_gate")
    close(ff.swiglu.w_up.grad, leaves["w_up"].grad, BWD, "dw_up")
    close(ff.w_down.weight.grad, leaves["w_down"].grad, BWD, "dw_down")


def test_chained_stack_equals_unchained_stack():
    """A stack of blocks wired with chain_block_norms gives bit-identical results to the same stack without the wiring (the
    chained norm is the same kernel on the same bytes), in inference and under autograd."""
    torch.manual_seed(3)
    hidden, inter, tokens, nblocks = 256, 688, 200, 3

    class Blk(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.norm1, self.norm2 = L.LLAMARMSNorm(hidden, 1e-5), L.LLAMARMSNorm(hidden, 1e-5)
            self.ff = L.FusedFeedforward(hidden, inter)
            self.mix = torch.nn.Linear(hidden, hidden, bias=False)       # stand-in for attention (not on this path)

        def forward(self, hs):
            attn_out = self.mix(self.norm1(hs))
            return L.block_tail(self.norm2, self.ff, attn_out, hs, next_norm=getattr(self, "_l32_next_norm", None))

    blocks = torch.nn.ModuleList([Blk() for _ in range(nblocks)]).to(DEV, torch.bfloat16)
    final = L.LLAMARMSNorm(hidden, 1e-5).to(DEV, torch.bfloat16)
    for m in list(blocks.modules()) + [final]:
        if isinstance(m, L.LLAMARMSNorm):
            with torch.no_grad():
                m.weight.add_(0.1 * torch.randn_like(m.weight))
    x = torch.randn(tokens, hidden, device=DEV).bfloat16()

    def run(x_in):
        hs = x_in
        for b in blocks:
            hs = b(hs)
        return final(hs)

    results = {}
    for mode in ("plain", "chained"):
        if mode == "chained":
            L.chain_block_norms(blocks, final)
        with torch.no_grad():
            y_inf = run(x)
        xin = x.clone().requires_grad_(True)
        for p in list(blocks.parameters()) + list(final.parameters()):
            p.grad = None
        run(xin).backward(torch.ones_like(x))
        results[mode] = (y_inf, xin.grad.clone(), blocks[0].ff.swiglu.w_gate.grad.clone(), final.weight.grad.clone())
    for a, b, what in zip(results["plain"], results["chained"], ("y", "dx", "dw_gate[0]", "d final gamma")):
        if what == "y":
            assert torch.equal(a, b), what
        else:
            close(b, a, BWD, what)      # same math; the fused addends round once instead of twice (bf16, three blocks deep)


@pytest.mark.parametrize("tokens,in_f,out_f,rank,p_drop,bias", [
    (300, 256, 512, 16, 0.0, False), (1030, 512, 256, 8, 0.0, True), (256, 4096, 1024, 16, 0.05, False), (77, 256, 256, 64, 0.25, False)])
def test_linear_lora_fused_any_projection(tokens, in_f, out_f, rank, p_drop, bias):
    """Linear_LORA on any projection (q/k/v/out of the attention as well as w_down): base GEMM + rank-r adapter as ONE
    kernel with two accumulation phases, LoRA dropout included, against autograd over Model/model.py:120-121."""
    g = torch.Generator().manual_seed(tokens + rank)
    bf = O.bf16_representable
    x32, dy32 = bf(torch.randn(tokens, in_f, generator=g)), bf(torch.randn(tokens, out_f, generator=g))
    w32 = bf((torch.rand(out_f, in_f, generator=g) * 2 - 1) / in_f ** 0.5)
    a32 = bf(torch.randn(rank, in_f, generator=g) / in_f ** 0.5)
    b32 = bf(0.05 * torch.randn(out_f, rank, generator=g))
    bias32 = bf(0.1 * torch.randn(out_f, generator=g)) if bias else None
    alpha = 32.0
    lo = L.Linear_LORA(in_f, out_f, rank=rank, alpha=alpha, dropout=p_drop).to(DEV, torch.bfloat16)
    if bias:
        lo.linear.bias = torch.nn.Parameter(dev(bias32), requires_grad=False)
    with torch.no_grad():
        lo.linear.weight.copy_(dev(w32)); lo.lora_a.weight.copy_(dev(a32)); lo.lora_b.weight.copy_(dev(b32))
    lo.train()
    x = dev(x32).requires_grad_(True)
    torch.manual_seed(99)
    y = lo(x)
    y.backward(dev(dy32))
    # the mask the layer drew: same generator state, same shape / dtype / device
    mask = torch.ones(tokens, in_f)
    if p_drop > 0:
        torch.manual_seed(99)
        mask = (torch.empty_like(x).bernoulli_(1.0 - p_drop).float() / (1.0 - p_drop)).cpu()
        assert 0.5 * p_drop < (mask == 0).float().mean().item() < 1.5 * p_drop + 0.02
    xs, As, Bs = x32.clone().requires_grad_(True), a32.clone().requires_grad_(True), b32.clone().requires_grad_(True)
    yr = torch.nn.functional.linear(xs, w32, bias32) + (alpha / rank) * torch.nn.functional.linear(
        torch.nn.functional.linear(xs * mask, As), Bs)
    yr.backward(dy32)
    tol_f, tol_b = (1e-2, 2.0 ** -6), (1.5e-2, 2.0 ** -5)
    close(y, yr, tol_f, "lora linear y")
    close(x.grad, xs.grad, tol_b, "dx")
    close(lo.lora_a.weight.grad, As.grad, tol_b, "dlora_a")
    close(lo.lora_b.weight.grad, Bs.grad, tol_b, "dlora_b")
    assert lo.linear.weight.grad is None
    # eval mode: no dropout, no autograd -> inference path of the same kernels
    lo.eval()
    with torch.no_grad():
        ye = lo(dev(x32))
    close(ye, torch.nn.functional.linear(x32, w32, bias32) + (alpha / rank) * (x32 @ a32.t()) @ b32.t(), tol_f, "eval y")


# ----------------------------------------------------------------------------------------------- f4: lm_head + shifted CE
def test_lm_head_ce_vs_reference_outputs():
    """Fused lm_head + shifted cross entropy against the reference's own logits and loss (fixture from Model/model.py:429-438)."""
    g = load_golden("lm_head_ce_cfg1.npz")
    head = torch.nn.Linear(256, 512, bias=False).to(DEV, torch.bfloat16)
    with torch.no_grad():
        head.weight.copy_(dev(g["lm_head_weight"]))
    hs = dev(g["hidden_states"])
    labels = g["labels"].to(DEV)
    logits, loss = L.lm_head_loss(head, hs, labels, int(g["ignore_index"]))
    close(logits, g["logits"], FWD, "logits")
    assert abs(loss.item() - g["loss"]) <= 5e-3 * abs(g["loss"]), (loss.item(), g["loss"])
    # without labels: plain logits, no loss
    lg2, none = L.lm_head_loss(head, hs, None)
    assert none is None and lg2.shape == logits.shape


@pytest.mark.parametrize("batch,seq,hidden,vocab,frac_ignored", [(2, 160, 256, 1000, 0.2), (3, 100, 512, 2056, 0.0), (1, 700, 384, 520, 0.9),
                                                                 (2, 64, 256, 128256, 0.1)])
def test_lm_head_ce_forward_backward_vs_oracle(batch, seq, hidden, vocab, frac_ignored):
    """Loss, logits, d_hidden and d_weight of the fused head against autograd over the reference's expressions; vocabularies
    that are not a multiple of the 256-column tile, rows without a target, and the real 128 256-entry vocabulary."""
    gen = torch.Generator().manual_seed(vocab + seq)
    bf = O.bf16_representable
    hs32 = bf(torch.randn(batch, seq, hidden, generator=gen))
    w32 = bf(torch.randn(vocab, hidden, generator=gen) / hidden ** 0.5)
    labels = torch.randint(0, vocab, (batch, seq), generator=gen)
    labels[torch.rand(batch, seq, generator=gen) < frac_ignored] = -100
    labels[0, 1] = 7                                    # at least one valid target
    head = torch.nn.Linear(hidden, vocab, bias=False).to(DEV, torch.bfloat16)
    with torch.no_grad():
        head.weight.copy_(dev(w32))
    hs = dev(hs32).requires_grad_(True)
    logits, loss = L.lm_head_loss(head, hs, labels.to(DEV))
    (loss * 3.0).backward()                              # a non-unit upstream gradient (read on the device, no sync)
    hr, wr = hs32.clone().requires_grad_(True), w32.clone().requires_grad_(True)
    logits_r, loss_r = O.lm_head_shifted_ce(hr, wr, labels)
    (loss_r * 3.0).backward()
    close(logits, logits_r, FWD, "logits")
    assert abs(loss.item() - loss_r.item()) <= 5e-3 * abs(loss_r.item()), (loss.item(), loss_r.item())
    close(hs.grad, hr.grad, (2e-2, 2.0 ** -5), "d_hidden")
    close(head.weight.grad, wr.grad, (2e-2, 2.0 ** -5), "d_lm_head_weight")


# ----------------------------------------------------------------------------------------------- f3: attention + KV cache
class _Cfg:
    def __init__(self, hidden, heads, kv, rope_base=500000.0):
        self.hidden_size, self.n_heads, self.n_kv_groups, self.rope_base = hidden, heads, kv, rope_base


def _gqa_module(hidden, heads, kv, wq, wk, wv, wo, rope_base=500000.0, layer_idx=0):
    att = L.GroupQueryAttention(_Cfg(hidden, heads, kv, rope_base), layer_idx=layer_idx).to(DEV, torch.bfloat16).eval()
    with torch.no_grad():
        att.W_query.weight.copy_(dev(wq)); att.W_key.weight.copy_(dev(wk)); att.W_value.weight.copy_(dev(wv))
        att.out_proj.weight.copy_(dev(wo))
    return att


@pytest.mark.parametrize("tag", ["d64", "d128"])
def test_attention_vs_reference_outputs(tag):
    """GroupQueryAttention drop-in (RoPE + preallocated KV cache + flash-style tcgen05 attention) against the reference's own
    module outputs: padded prefill, then a KV-cached decode step with explicit position_ids; cache contents included."""
    g = load_golden(f"attention_{tag}.npz")
    heads, kv = int(g["n_heads"]), int(g["n_kv"])
    b, t, hidden = g["x"].shape
    att = _gqa_module(hidden, heads, kv, g["wq"], g["wk"], g["wv"], g["wo"], float(g["rope_base"]))
    assert sorted(att.state_dict().keys()) == list(g["state_dict_keys"])
    mask4d = O.causal_padding_mask(g["mask2d"], t).to(DEV, torch.bfloat16)
    cache = L.KVCache()
    with torch.no_grad():
        y = att(dev(g["x"]), attention_mask=mask4d, position_ids=g["position_ids"].to(DEV), kv_cache=cache)
        assert cache.num_items() == t
        y1 = att(dev(g["x_decode"]), attention_mask=torch.zeros(b, 1, 1, 1, device=DEV, dtype=torch.bfloat16),
                 position_ids=g["position_ids_decode"].to(DEV), kv_cache=cache)
    close(y, g["y_prefill"], FWD, "prefill output")
    close(y1, g["y_decode"], FWD, "decode output")
    assert cache.num_items() == t + 1
    close(cache.key_cache[0], g["cache_k"], FWD, "cache keys (rotated)")
    close(cache.value_cache[0], g["cache_v"], FWD, "cache values")


@pytest.mark.parametrize("tokens", [1, 5, 64, 128, 129, 300, 2048])
@pytest.mark.parametrize("outs", [(4096, 1024, 1024), (512, 128, 128), (1024, 1024), (256,)])
def test_linear_group_forward_vs_oracle(tokens, outs):
    """l32_linear_group_forward (the W_query / W_key / W_value projections of GroupQueryAttention.forward, reference
    Model/model.py:231-233, as one launch): small-M weight-streaming kernel over the row blocks of all weights (tokens <= 128)
    and the grouped tcgen05 GEMM (above), against the three F.linear calls in fp32 and against the single-projection call."""
    from llama32_b200 import ops
    k = 4096 if outs[0] == 4096 else 512
    gen = torch.Generator().manual_seed(tokens * 7 + len(outs))
    rep = lambda v: v.to(torch.bfloat16).float()
    a = rep(torch.randn(tokens, k, generator=gen))
    ws = [rep((torch.rand(o, k, generator=gen) * 2 - 1) / k ** 0.5) for o in outs]
    ys = ops.linear_group_forward(dev(a), [dev(w) for w in ws])
    assert len(ys) == len(outs)
    for y, w in zip(ys, ws):
        close(y, a @ w.t(), FWD, f"grouped projection {tuple(w.shape)} at {tokens} tokens")
        # (not bit-equal: the grouped grid may split K differently, i.e. sum the fp32 partials in another order)
        close(y, ops.linear_forward(dev(a), dev(w)).float(), (4e-3, 2.0 ** -7), "grouped launch vs the single projection")


@pytest.mark.parametrize("b,t,heads,kv,d,causal,pad", [(1, 128, 1, 1, 128, True, 0), (1, 128, 1, 1, 128, True, 5), (1, 128, 1, 1, 128, True, 37),
                                                       (1, 128, 1, 1, 128, True, 64), (1, 128, 1, 1, 128, True, 70),
                                                       (2, 300, 4, 2, 128, True, 37), (2, 300, 4, 2, 64, True, 37),
                                                       (2, 512, 8, 2, 128, False, 0), (1, 2048, 8, 2, 128, True, 100)])
def test_attention_kernel_key_padding_and_warp_skew(b, t, heads, kv, d, causal, pad):
    """The attention kernel alone (l32_gqa_attention_forward) with a key-padding vector against the softmax written out in
    fp32.  With padding bytes the four softmax warps of a CTA take very different times per tile (a warp whose rows see no key
    of a tile skips the mask loop), which is what exposes ordering mistakes between the warps, the MMA issuer and the
    barriers' phases: left padding shorter / longer than a warp, a whole key tile, and more than a tile; dead rows must give
    exact zeros."""
    from llama32_b200 import ops
    gen = torch.Generator().manual_seed(b * 131 + t + pad)
    rep = lambda v: v.to(torch.bfloat16).float()
    q, k, v = rep(torch.randn(b, t, heads * d, generator=gen)), rep(torch.randn(b, kv, t + 8, d, generator=gen)), rep(torch.randn(b, kv, t + 8, d, generator=gen))
    keep = torch.ones(b, t, dtype=torch.uint8)
    keep[0, :pad] = 0
    y = ops.gqa_attention_forward(dev(q), dev(k), dev(v), t, 0, causal=causal, key_keep=keep.to(DEV))
    q4 = q.view(b, t, heads, d).transpose(1, 2)
    k4, v4 = k[:, :, :t].repeat_interleave(heads // kv, 1), v[:, :, :t].repeat_interleave(heads // kv, 1)
    sc = q4 @ k4.transpose(-1, -2) / d ** 0.5
    m = torch.ones(t, t, dtype=torch.bool)
    if causal:
        m &= torch.arange(t)[None] <= torch.arange(t)[:, None]
    m = m[None, None] & keep[:, None, None, :].bool()
    pr = torch.nan_to_num(torch.softmax(sc.masked_fill(~m, float("-inf")), -1), nan=0.0)     # a row without keys: zeros
    ref = (pr @ v4).transpose(1, 2).reshape(b, t, heads * d)
    close(y, ref, FWD, "attention with key padding")
    dead = ~m.any(-1)[:, 0]                                                       # [b, t]: rows that see no key at all
    assert (y.float().cpu()[dead] == 0).all(), "rows without any visible key must be exact zeros"


@pytest.mark.parametrize("b,t,hidden,heads,kv,steps", [(2, 300, 1024, 8, 2, 3), (1, 512, 4096, 32, 8, 2), (3, 129, 512, 8, 8, 1),
                                                     (2, 64, 256, 4, 1, 2)])
def test_attention_shapes_vs_oracle(b, t, hidden, heads, kv, steps):
    """More geometries (the 11B one: 32 query / 8 KV heads of 128) against the oracle: causal prefill with left-over tiles,
    a padded batch, several decode steps growing the preallocated cache, and the no-mask (non-causal) call."""
    gen = torch.Generator().manual_seed(b * 1000 + t)
    bf = O.bf16_representable
    d = hidden // heads
    mk = lambda o, i: bf((torch.rand(o, i, generator=gen) * 2 - 1) / i ** 0.5)
    wq, wk, wv, wo = mk(heads * d, hidden), mk(kv * d, hidden), mk(kv * d, hidden), mk(hidden, heads * d)
    att = _gqa_module(hidden, heads, kv, wq, wk, wv, wo)
    x = bf(torch.randn(b, t, hidden, generator=gen))
    mask2d = torch.ones(b, t)
    if b > 1:
        mask2d[-1, t - 5:] = 0
    pos = torch.arange(t)[None].expand(b, -1).contiguous()
    cache = L.KVCache(capacity=t + 1)                       # forces one re-allocation when steps > 1
    with torch.no_grad():
        y = att(dev(x), attention_mask=O.causal_padding_mask(mask2d, t).to(DEV, torch.bfloat16), position_ids=pos.to(DEV), kv_cache=cache)
    yr, kr, vr = O.gqa_attention(x, wq, wk, wv, wo, heads, kv, pos, O.causal_padding_mask(mask2d, t))
    close(y, yr, FWD, "prefill")
    for sidx in range(steps):
        x1 = bf(torch.randn(b, 1, hidden, generator=gen))
        p1 = torch.full((b, 1), t + sidx, dtype=torch.long)
        with torch.no_grad():
            y1 = att(dev(x1), attention_mask=torch.zeros(b, 1, 1, 1, device=DEV, dtype=torch.bfloat16), position_ids=p1.to(DEV), kv_cache=cache)
        y1r, kr, vr = O.gqa_attention(x1, wq, wk, wv, wo, heads, kv, p1, torch.zeros(b, 1, 1, 1), past_k=kr, past_v=vr)
        close(y1, y1r, FWD, f"decode step {sidx}")
    assert cache.num_items() == t + steps
    # attention_mask=None: no masking at all (the reference adds nothing), no cache
    with torch.no_grad():
        yn = att(dev(x), attention_mask=None, position_ids=pos.to(DEV), kv_cache=None)
    ynr, _, _ = O.gqa_attention(x, wq, wk, wv, wo, heads, kv, pos, None)
    close(yn, ynr, FWD, "no mask")


def test_odd_feature_sizes_take_the_reference_expressions():
    """hidden / inter that are not multiples of 8 (no 16-byte rows for TMA): the reference's live path accepts anything, so the
    drop-in modules evaluate the reference's F.linear expressions there instead of raising (smaller #14)."""
    torch.manual_seed(0)
    ff = L.FusedFeedforward(60, 100).to(DEV, torch.bfloat16)
    x = torch.randn(3, 7, 60, device=DEV).bfloat16().requires_grad_(True)
    y = ff(x)
    y.sum().backward()
    ref = torch.nn.functional.linear(torch.nn.functional.silu(torch.nn.functional.linear(x, ff.swiglu.w_gate)) *
                                     torch.nn.functional.linear(x, ff.swiglu.w_up), ff.w_down.weight)
    assert torch.equal(y, ref) and x.grad is not None and ff.swiglu.w_gate.grad is not None


def _layer(hidden, heads, kv, inter, dtype=torch.bfloat16, seed=0):
    """One decoder layer wired like the reference's TransformerBlock (Model/model.py:257-273) from the drop-in modules."""
    torch.manual_seed(seed)

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.att = L.GroupQueryAttention(_Cfg(hidden, heads, kv), layer_idx=0)
            self.norm1, self.norm2 = L.LLAMARMSNorm(hidden, 1e-5), L.LLAMARMSNorm(hidden, 1e-5)
            self.ff = L.FusedFeedforward(hidden, inter)

        def forward(self, hs, attention_mask=None, position_ids=None, kv_cache=None):
            attn_out = self.att(self.norm1(hs), attention_mask=attention_mask, position_ids=position_ids, kv_cache=kv_cache)
            return L.block_tail(self.norm2, self.ff, attn_out, hs)
    return Block().to(DEV, dtype).eval()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_whole_layer_prefill_and_graph_captured_decode(dtype):
    """norm1 -> attention (RoPE, preallocated KV cache, flash-style kernel) -> fused block tail as ONE layer, in bf16 and fp16:
    prefill against the oracle, then decode steps replayed from a CUDA graph (nothing in the C ABI allocates or synchronises;
    the cache is preallocated, so the captured addresses stay valid) against the oracle and against eager execution."""
    hidden, heads, kv, inter, b, t = 512, 4, 2, 1408, 2, 96
    blk = _layer(hidden, heads, kv, inter, dtype)
    rep = (lambda v: v.to(dtype).float())
    f = lambda m: m.weight.detach().float().cpu()
    gen = torch.Generator().manual_seed(1)
    x = rep(torch.randn(b, t, hidden, generator=gen))
    pos = torch.arange(t)[None].expand(b, -1).contiguous()
    mask = O.causal_padding_mask(torch.ones(b, t), t)

    def oracle(xin, posn, m4, pk, pv):
        n1 = O.add_rmsnorm(xin, f(blk.norm1), 1e-5)
        a, k, v = O.gqa_attention(n1, f(blk.att.W_query), f(blk.att.W_key), f(blk.att.W_value), f(blk.att.out_proj), heads, kv, posn, m4,
                                  past_k=pk, past_v=pv)
        _, _, out = O.block_hot_path(a, xin, f(blk.norm2), 1e-5, blk.ff.swiglu.w_gate.detach().float().cpu(),
                                     blk.ff.swiglu.w_up.detach().float().cpu(), f(blk.ff.w_down))
        return out, k, v

    cache = L.KVCache(capacity=t + 8)
    with torch.no_grad():
        y = blk(x.to(DEV, dtype), attention_mask=mask.to(DEV, dtype), position_ids=pos.to(DEV), kv_cache=cache)
    yr, kr, vr = oracle(x, pos, mask, None, None)
    tol = (1.5e-2, 2.0 ** -5)                                   # a whole layer deep in 16-bit storage
    close(y, yr, tol, "layer prefill")
    # ---- decode: static input buffers, one captured step, replayed with new contents
    x1_buf = torch.zeros(b, 1, hidden, device=DEV, dtype=dtype)
    p1_buf = torch.zeros(b, 1, device=DEV, dtype=torch.long)
    zero_mask = torch.zeros(b, 1, 1, 1, device=DEV, dtype=dtype)
    steps = 3
    xs = [rep(torch.randn(b, 1, hidden, generator=gen)) for _ in range(steps)]
    # eager reference run on a copy of the cache state
    import copy
    cache_eager = copy.deepcopy(cache)
    eager = []
    with torch.no_grad():
        for i in range(steps):
            eager.append(blk(xs[i].to(DEV, dtype), attention_mask=zero_mask, position_ids=torch.full((b, 1), t + i, device=DEV),
                             kv_cache=cache_eager).clone())
    # captured: every step appends at a different position, so the step is captured per position (one graph per cache length,
    # as a serving loop would do per bucket); the point is that capture works and replays to the same bits
    outs = []
    with torch.no_grad():
        for i in range(steps):
            x1_buf.copy_(xs[i].to(DEV, dtype)); p1_buf.fill_(t + i)
            torch.cuda.synchronize()
            gph = torch.cuda.CUDAGraph()
            cache_len = cache.num_items()
            with torch.cuda.graph(gph):
                y1 = blk(x1_buf, attention_mask=zero_mask, position_ids=p1_buf, kv_cache=cache)
            cache._len[0] = cache_len                             # capture does not execute: rewind, then replay for real
            cache.advance(0, 0)
            gph.replay()
            cache.advance(0, 1)
            torch.cuda.synchronize()
            outs.append(y1.clone())
    for i in range(steps):
        y1r, kr, vr = oracle(xs[i], torch.full((b, 1), t + i, dtype=torch.long), torch.zeros(b, 1, 1, 1), kr, vr)
        close(outs[i], y1r, tol, f"layer decode step {i} (graph replay)")
        assert torch.equal(outs[i], eager[i]), f"graph replay differs from eager execution at step {i}"
