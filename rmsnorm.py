"""`rmsnorm` extension entry points (same module name and callables as the reference's CUDAExtension
`rmsnorm`, reference setup.py:11-18 / Tools/rmsnorm/rmsnorm.cu:63-67), bound to the sm_100a C-ABI.

    forward(input, weight, residual, eps) -> [output, rms]
    backward(grad_out, input, weight, rms) -> [d_input, d_weight]

Differences from the reference kernels, all deliberate:
  * bf16 as well as fp16 (the reference hard-codes at::Half, rmsnorm.cu:22-27);
  * the residual add is done in fp32 registers; `residual` is updated in place to x + residual exactly as
    the reference kernel does (rmsnorm.cuh:50-55) because callers of THIS raw ABI may rely on it -- the
    shipped autograd wrapper (llama32_b200.RMSNormFunction) never mutates caller tensors;
  * `backward` expects `input` = the tensor that was normalised (x + residual), which is what makes the
    reference's formula correct.
"""
import torch

from llama32_b200 import ops as _ops


def forward(input, weight, residual, eps):
    res = residual
    if res is not None and not res.is_contiguous():
        raise RuntimeError("rmsnorm.forward: residual must be contiguous (it is updated in place)")
    y, rms, _ = _ops.add_rmsnorm_forward(input, weight, res, eps, want_rms=True,
                                         h_out=None if res is None else res.view(input.shape))
    return [y, rms]


def backward(grad_out, input, weight, rms):
    dx, dw = _ops.rmsnorm_backward(grad_out, input, weight, rms, want_dweight=True)
    return [dx, dw]


__all__ = ["forward", "backward"]
