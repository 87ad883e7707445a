"""`swiglu_fused` extension entry points (same module name and callables as the reference's CUDAExtension
`swiglu_fused`, reference setup.py:21-41 / Tools/swiglu/swiglu_binding.cpp:5-32), bound to the sm_100a C-ABI.

    forward(x, w_gate, w_up, b_gate=None, b_up=None) -> [output, gate_cache, up_cache]
    backward(grad_output, x, w_gate, w_up, gate_cache, up_cache) -> (grad_x, grad_w_gate, grad_w_up)
    forward_down(x, w_gate, w_up, w_down, b_gate=None, b_up=None, b_down=None) -> output

Weights use the Python/HF layout ([inter, hidden] / [hidden, inter], Tools/swiglu/FusedSwiglu.py:63-64), not
the transposed layout the reference .cu assumed (swiglu.cu:89-90).  `None` biases are accepted (the reference
binding could not accept them).
"""
from llama32_b200 import ops as _ops


def forward(x, w_gate, w_up, b_gate=None, b_up=None):
    act, gate, up = _ops.swiglu_forward(x, w_gate, w_up, b_gate, b_up, want_cache=True)
    return [act, gate, up]


def backward(grad_output, x, w_gate, w_up, gate_cache, up_cache):
    dx, dwg, dwu, _, _ = _ops.swiglu_backward(grad_output, x, w_gate, w_up, gate_cache, up_cache)
    return dx, dwg, dwu


def forward_down(x, w_gate, w_up, w_down, b_gate=None, b_up=None, b_down=None):
    y, _, _ = _ops.ffn_forward(x, w_gate, w_up, w_down, b_gate, b_up, b_down, want_cache=False)
    return y


__all__ = ["forward", "backward", "forward_down"]
