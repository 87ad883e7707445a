"""The zero-code-change route on the GPU: the UNPATCHED reference wrappers find the modules
`rmsnorm` and `swiglu_fused` (reference Model/model.py:8 `find_spec("rmsnorm")`, Tools/swiglu/FusedSwiglu.py:7
`import swiglu_fused as swiglu`) and call them in their own way.  The reference tree does not exist on the GPU box, so
the wrappers' call sequences are replayed here step by step, argument for argument:

  RMSNormFunction.forward   (Model/model.py:137-146): x.contiguous(); weight.contiguous() -- an fp32 Parameter next to 16-bit
                            activations; residual = zeros_like(x) when None; `output, rms = rmsnorm.forward(x, weight,
                            residual, eps)`; saves (x, weight, rms).
  RMSNormFunction.backward  (Model/model.py:148-155): `d_x, d_weight = rmsnorm.backward(grad_output, x, weight, rms)`;
                            casts d_weight to weight.dtype; returns (d_x, d_weight, None, d_x).
  SwiGLUFunction.forward    (Tools/swiglu/FusedSwiglu.py:22-29): `swiglu.forward(x, w_gate, w_up, b_gate, b_up)` with five
                            positional arguments, biases None, x 3-D; saves the two caches.
  SwiGLUFunction.backward   (FusedSwiglu.py:32-40): `swiglu.backward(grad_output, x, w_gate, w_up, gate_cache, up_cache)`
                            -> 3-tuple; returns (grad_x, grad_w_gate, grad_w_up, None, None).
  swiglu.forward_down       (Tools/swiglu/swiglu_binding.cpp:24-31): 4 positional tensors + default biases.
Results are checked against the CPU oracle with the tolerances of tests/test_gpu_parity.py.
"""
import importlib
import importlib.util

import pytest
import torch

from oracle import ffn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
FWD = (1e-2, 2.0 ** -6)
BWD = (1e-2, 2.0 ** -5)


def close(got, ref, tol, what):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    r, m = O.rel_l2(got, ref), O.max_abs_over_max_ref(got, ref)
    assert r <= tol[0] and m <= tol[1], f"{what}: rel-L2 {r:.3e}, max-abs/max|ref| {m:.3e}"


def _ext(name):
    # exactly what the reference does to detect / load the extensions
    assert importlib.util.find_spec(name) is not None, f"module `{name}` is not importable from the repository root"
    return importlib.import_module(name)


class _RefStyleRMSNorm(torch.autograd.Function):
    """The call sequence of the reference's RMSNormFunction (Model/model.py:135-155), replayed against our `rmsnorm`."""

    @staticmethod
    def forward(ctx, x, weight, eps, residual=None):
        rmsnorm = _ext("rmsnorm")
        x = x.contiguous()
        weight = weight.contiguous()
        if residual is None:
            residual = torch.zeros_like(x)
        residual = residual.contiguous()
        output, rms = rmsnorm.forward(x, weight, residual, eps)
        ctx.save_for_backward(x, weight, rms)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        rmsnorm = _ext("rmsnorm")
        x, weight, rms = ctx.saved_tensors
        d_x, d_weight = rmsnorm.backward(grad_output.contiguous(), x, weight, rms)
        if d_weight.dtype != weight.dtype:
            d_weight = d_weight.to(weight.dtype)
        # the reference returns `d_x` for the residual slot unconditionally (Model/model.py:155); autograd rejects a gradient
        # for a None input, so with residual=None (the only case replayed through backward) the slot must be None
        return d_x, d_weight, None, (d_x if ctx.needs_input_grad[3] else None)


class _RefStyleSwiGLU(torch.autograd.Function):
    """The call sequence of the reference's SwiGLUFunction (Tools/swiglu/FusedSwiglu.py:14-40) against our `swiglu_fused`."""

    @staticmethod
    def forward(ctx, x, w_gate, w_up, b_gate=None, b_up=None):
        swiglu = _ext("swiglu_fused")
        x, w_gate, w_up = x.contiguous(), w_gate.contiguous(), w_up.contiguous()
        output, gate_cache, up_cache = swiglu.forward(x, w_gate, w_up, b_gate, b_up)
        ctx.save_for_backward(x, w_gate, w_up, gate_cache, up_cache)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        swiglu = _ext("swiglu_fused")
        x, w_gate, w_up, gate_cache, up_cache = ctx.saved_tensors
        grad_x, grad_w_gate, grad_w_up = swiglu.backward(grad_output.contiguous(), x, w_gate, w_up, gate_cache, up_cache)
        return grad_x, grad_w_gate, grad_w_up, None, None


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_unpatched_rmsnorm_wrapper_sequence(dtype):
    torch.manual_seed(0)
    B, S, H, eps = 2, 160, 256, 1e-5
    rep = (lambda t: t.to(torch.bfloat16).float()) if dtype == torch.bfloat16 else (lambda t: t.half().float())
    x32, g32 = rep(torch.randn(B, S, H)), rep(torch.randn(B, S, H))
    weight = torch.nn.Parameter(rep(1 + 0.1 * torch.randn(H)).to(DEV))          # fp32 Parameter, as LLAMARMSNorm creates it
    assert weight.dtype == torch.float32
    # norm1-style call: no residual -> the wrapper passes zeros_like(x)
    x = x32.to(DEV, dtype).requires_grad_(True)
    y = _RefStyleRMSNorm.apply(x, weight, eps, None)
    assert y.dtype == dtype and y.shape == x.shape
    y.backward(g32.to(DEV, dtype))
    yr, dxr, dwr, _ = O.add_rmsnorm_grads(x32, weight.detach().float().cpu(), eps, None, g32)
    close(y, yr, FWD, "y (no residual)")
    close(x.grad, dxr, BWD, "dx")
    close(weight.grad, dwr, BWD, "dweight (cast back to the fp32 Parameter)")
    assert weight.grad.dtype == torch.float32
    # norm2-style call: with a residual; the raw ABI updates it in place to x + residual like the reference kernel
    r32 = rep(torch.randn(B, S, H))
    res = r32.to(DEV, dtype)
    with torch.no_grad():
        y2 = _RefStyleRMSNorm.apply(x.detach(), weight, eps, res)
    close(y2, O.add_rmsnorm(x32, weight.detach().float().cpu(), eps, r32), FWD, "y (residual)")
    close(res, x32 + r32, (4e-3, 2.0 ** -7), "residual := x + residual (in place, reference rmsnorm.cuh:50-55)")


def test_unpatched_swiglu_wrapper_sequence():
    B, S, H, I = 2, 96, 256, 688                      # config 1's text shapes
    s = O.synthetic_ffn(B * S, H, I, seed=4)
    bf = lambda t: t.to(DEV, torch.bfloat16)
    x = bf(s["x"]).view(B, S, H).requires_grad_(True)            # 3-D, as swiglu.cu:284-286 requires
    wg = bf(s["w_gate"]).requires_grad_(True)
    wu = bf(s["w_up"]).requires_grad_(True)
    act = _RefStyleSwiGLU.apply(x, wg, wu, None, None)
    assert act.shape == (B, S, I)
    d_act = torch.randn(B, S, I).bfloat16()
    act.backward(d_act.to(DEV))
    ref = O.swiglu_grads(s["x"].view(B, S, H), s["w_gate"], s["w_up"], d_act.float())
    close(act, ref["act"], FWD, "act")
    close(x.grad, ref["dx"], BWD, "grad_x")
    close(wg.grad, ref["dw_gate"], BWD, "grad_w_gate")
    close(wu.grad, ref["dw_up"], BWD, "grad_w_up")
    # forward_down with the defaults of the binding (swiglu_binding.cpp:24-31)
    swiglu = _ext("swiglu_fused")
    y = swiglu.forward_down(x.detach(), wg.detach(), wu.detach(), bf(s["w_down"]))
    close(y, O.feedforward(s["x"], s["w_gate"], s["w_up"], s["w_down"]).view(B, S, H), FWD, "forward_down")
