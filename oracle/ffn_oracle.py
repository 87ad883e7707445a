"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Add-RMSNorm + SwiGLU feed-forward hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker or the reported CPU baseline; the product path (llama32_b200/*) never does.

It restates, in plain fp32/fp64 PyTorch on the CPU, the reference's LIVE path for this hot path -- the
PyTorch expressions the reference itself executes whenever its CUDA extensions are absent or the input is
not a 16-bit CUDA tensor (which, for the FFN, is always: SURVEY.md section 0.1):
    add_rmsnorm   <- LLAMARMSNorm.forward fallback branch      reference Model/model.py:166-171
    swiglu        <- SwiGLUFunction.forward fallback branch    reference Tools/swiglu/FusedSwiglu.py:17-20
    feedforward   <- FusedFeedforward.forward                  reference Model/model.py:216-217
    block_hot_path<- TransformerBlock.forward lines norm2/ff   reference Model/model.py:270-273
    linear_lora   <- Linear_LORA.forward                       reference Model/model.py:120-121
    lm_head_shifted_ce <- MllamaForConditionalGeneration.forward tail   reference Model/model.py:429-438
    gqa_attention (+ rope_cos_sin, rotate_half, causal_padding_mask) <- GroupQueryAttention.forward, LLAMARotaryEmbedding,
                  KVCache.update, Llama3Model._prepare_attention_mask   reference Model/model.py:12-29, 176-198, 220-254, 304-319
The arithmetic itself lives in the third-party dependency torch (reference setup.py:47 `torch>=2.0.0`;
installed here: 2.11.0+cu128): F.linear, F.silu, rsqrt, mean.  Gradients: the reference's own backward
functions cannot run on any path, so the gradient oracle is autograd over these
same expressions, cross-checked below against closed-form derivatives in fp64.

Pinning: the reference ships no tests and no golden vectors, so the oracle is pinned
against outputs of the reference itself, generated in the build container by oracle/make_golden.py (which
imports /root/reference) and committed under tests/golden/; tests/test_oracle.py checks them bit-for-bit.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ forward
def add_rmsnorm(x, weight, eps, residual=None):
    """reference Model/model.py:166-171 (same op order: add, pow(2).mean, rsqrt(var + eps), * weight)."""
    if residual is not None:
        x = x + residual
    variance = x.pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(variance + eps)
    return x * weight


def rms_of(x, eps, residual=None):
    """sqrt(mean(h^2) + eps): the per-row statistic the reference CUDA kernel returns (rmsnorm.cuh:79-83)."""
    h = x if residual is None else x + residual
    return torch.sqrt(h.pow(2).mean(-1) + eps)


def swiglu(x, w_gate, w_up, b_gate=None, b_up=None):
    """reference Tools/swiglu/FusedSwiglu.py:18-20."""
    gate = F.linear(x, w_gate, b_gate)
    up = F.linear(x, w_up, b_up)
    return F.silu(gate) * up


def gate_up(x, w_gate, w_up, b_gate=None, b_up=None):
    """The two projections the reference extension caches for backward (swiglu.cu:93-98)."""
    return F.linear(x, w_gate, b_gate), F.linear(x, w_up, b_up)


def feedforward(x, w_gate, w_up, w_down, b_gate=None, b_up=None, b_down=None):
    """reference Model/model.py:217: w_down(swiglu(x)), w_down = nn.Linear(I, H) (model.py:214)."""
    return F.linear(swiglu(x, w_gate, w_up, b_gate, b_up), w_down, b_down)


def linear_lora(x, w, lora_a, lora_b, alpha, rank):
    """reference Model/model.py:120-121 with dropout p = 0 (eval / deterministic)."""
    return F.linear(x, w) + (alpha / rank) * F.linear(F.linear(x, lora_a), lora_b)


def block_hot_path(attn_out, hidden_states, norm2_weight, eps, w_gate, w_up, w_down):
    """reference Model/model.py:270-273: normed = norm2(attn_out, residual=hidden); return attn_out + ff(normed).
    (The skip connection is dropped from the block output -- SURVEY.md section 0.6; parity means keeping that.)"""
    normed = add_rmsnorm(attn_out, norm2_weight, eps, residual=hidden_states)
    ff_out = feedforward(normed, w_gate, w_up, w_down)
    return normed, ff_out, attn_out + ff_out


def rope_cos_sin(position_ids, head_dim, rope_base=500000.0):
    """reference Model/model.py:176-186 (LLAMARotaryEmbedding.forward): inv_freq = base^(-2i/d); emb = cat(freqs, freqs)."""
    inv_freq = 1.0 / (rope_base ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = (inv_freq[None, :, None].float().expand(position_ids.shape[0], -1, 1) @ position_ids[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    """reference Model/model.py:189-192."""
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def causal_padding_mask(mask2d, seq_len, dtype=torch.float32):
    """reference Model/model.py:304-319 (_prepare_attention_mask): triu(-inf, 1) + (1 - mask) * finfo.min, [B, 1, S, S]."""
    bsz = mask2d.shape[0]
    causal = torch.triu(torch.full((seq_len, seq_len), float("-inf"), dtype=dtype), diagonal=1)[None, None].expand(bsz, 1, seq_len, seq_len)
    padding = ((1.0 - mask2d.to(dtype)) * torch.finfo(dtype).min)[:, None, None, :].expand(bsz, 1, seq_len, seq_len)
    return causal + padding


def gqa_attention(hidden_states, wq, wk, wv, wo, n_heads, n_kv, position_ids, attention_mask=None, past_k=None, past_v=None,
                  rope_base=500000.0):
    """reference Model/model.py:238-253 (GroupQueryAttention.forward) with KVCache.update (:21-29) as an explicit concat.
    Returns (out, k_all, v_all) with k_all / v_all the cache contents [B, n_kv, len, d] after the call."""
    b, t, _ = hidden_states.shape
    d = wq.shape[0] // n_heads
    q = F.linear(hidden_states, wq).view(b, t, n_heads, d).transpose(1, 2)
    k = F.linear(hidden_states, wk).view(b, t, n_kv, d).transpose(1, 2)
    v = F.linear(hidden_states, wv).view(b, t, n_kv, d).transpose(1, 2)
    cos, sin = rope_cos_sin(position_ids, d, rope_base)
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    q, k = (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)
    if past_k is not None:
        k, v = torch.cat([past_k, k], dim=-2), torch.cat([past_v, v], dim=-2)
    g = n_heads // n_kv
    kr = k[:, :, None].expand(b, n_kv, g, k.shape[-2], d).reshape(b, n_heads, k.shape[-2], d)      # repeat_kv, model.py:124-132
    vr = v[:, :, None].expand(b, n_kv, g, v.shape[-2], d).reshape(b, n_heads, v.shape[-2], d)
    score = q @ kr.transpose(2, 3)
    if attention_mask is not None:
        score = score + attention_mask[:, :, :, : kr.shape[-2]]
    w = torch.softmax(score / (d ** 0.5), dim=-1)
    ctx = (w @ vr).transpose(1, 2).contiguous().reshape(b, t, -1)
    return F.linear(ctx, wo), k, v


def lm_head_shifted_ce(hidden_states, lm_head_weight, labels, ignore_index=-100):
    """reference Model/model.py:429-438: logits = lm_head(hidden_states) (nn.Linear, bias=False, model.py:354);
    shift_logits = logits[..., :-1, :]; shift_labels = labels[..., 1:]; CrossEntropyLoss(ignore_index) over the flattened rows.
    Returns (logits, loss)."""
    logits = F.linear(hidden_states, lm_head_weight)
    shift_logits = logits[..., :-1, :].contiguous()
    shift_labels = labels[..., 1:].contiguous()
    loss = torch.nn.CrossEntropyLoss(ignore_index=ignore_index)(shift_logits.view(-1, shift_logits.size(-1)),
                                                                shift_labels.view(-1))
    return logits, loss


# ------------------------------------------------------------------------------------------------ gradients
def add_rmsnorm_grads(x, weight, eps, residual, grad_out):
    """autograd over add_rmsnorm: returns (y, dx, dweight, dresidual|None)."""
    xs = x.detach().clone().requires_grad_(True)
    ws = weight.detach().clone().requires_grad_(True)
    rs = None if residual is None else residual.detach().clone().requires_grad_(True)
    y = add_rmsnorm(xs, ws, eps, rs)
    y.backward(grad_out)
    return y.detach(), xs.grad, ws.grad, (None if rs is None else rs.grad)


def add_rmsnorm_grads_closed_form(h, weight, eps, grad_out):
    """Closed form used by the CUDA kernel (same algebra as reference rmsnorm.cuh:124-152 with inp := h):
    rstd = rsqrt(mean(h^2)+eps); xhat = h*rstd; dx = rstd*(g*w - xhat*mean(g*w*xhat)); dw = sum_rows g*xhat."""
    rstd = torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + eps)
    xhat = h * rstd
    wdy = grad_out * weight
    dx = rstd * (wdy - xhat * (wdy * xhat).mean(-1, keepdim=True))
    dw = (grad_out * xhat).reshape(-1, h.shape[-1]).sum(0)
    return dx, dw


def feedforward_grads(x, w_gate, w_up, w_down, grad_out):
    """autograd over feedforward: returns dict(y, dx, dw_gate, dw_up, dw_down)."""
    xs, gs, us, ds = (t.detach().clone().requires_grad_(True) for t in (x, w_gate, w_up, w_down))
    y = feedforward(xs, gs, us, ds)
    y.backward(grad_out)
    return dict(y=y.detach(), dx=xs.grad, dw_gate=gs.grad, dw_up=us.grad, dw_down=ds.grad)


def swiglu_grads(x, w_gate, w_up, grad_act):
    """autograd over swiglu: returns dict(act, dx, dw_gate, dw_up)."""
    xs, gs, us = (t.detach().clone().requires_grad_(True) for t in (x, w_gate, w_up))
    act = swiglu(xs, gs, us)
    act.backward(grad_act)
    return dict(act=act.detach(), dx=xs.grad, dw_gate=gs.grad, dw_up=us.grad)


def swiglu_grads_closed_form(x, w_gate, w_up, grad_act):
    """Closed form the CUDA backward implements (math of the reference's unlaunched kernel swiglu.cu:204-221):
    s = sigmoid(g); d_gate = dA*u*s*(1+g*(1-s)); d_up = dA*g*s; dx = d_gate Wg + d_up Wu; dW = d^T x."""
    g, u = gate_up(x, w_gate, w_up)
    s = torch.sigmoid(g)
    d_gate = grad_act * u * (s * (1 + g * (1 - s)))
    d_up = grad_act * (g * s)
    x2 = x.reshape(-1, x.shape[-1])
    dg2, du2 = d_gate.reshape(-1, g.shape[-1]), d_up.reshape(-1, g.shape[-1])
    return dict(dx=d_gate @ w_gate + d_up @ w_up, dw_gate=dg2.t() @ x2, dw_up=du2.t() @ x2, d_gate=d_gate, d_up=d_up)


# ------------------------------------------------------------------------------------------------ helpers
def bf16_representable(t: torch.Tensor) -> torch.Tensor:
    """Round an fp32 tensor to the nearest bf16 value, kept in fp32 (both sides of a parity test then see
    identical inputs)."""
    return t.to(torch.bfloat16).to(torch.float32)


def fp16_representable(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float16).to(torch.float32)


def synthetic_ffn(tokens, hidden, inter, seed=0, device="cpu", gamma_noise=0.1):
    """Synthetic inputs of SURVEY.md section 8(d): x, residual, dY ~ N(0,1); w_gate, w_up ~ U(+-1/sqrt(H)) and
    w_down ~ U(+-1/sqrt(I)) (what the modules' kaiming_uniform(a=sqrt(5)) init produces); gamma = 1 + 0.1 N(0,1).
    All values bf16-representable fp32."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    u = lambda *s: torch.rand(*s, generator=g) * 2 - 1
    d = dict(
        x=r(tokens, hidden), residual=r(tokens, hidden), dy=r(tokens, hidden),
        w_gate=u(inter, hidden) / hidden ** 0.5, w_up=u(inter, hidden) / hidden ** 0.5,
        w_down=u(hidden, inter) / inter ** 0.5, gamma=1 + gamma_noise * r(hidden),
    )
    return {k: bf16_representable(v).to(device) for k, v in d.items()}


def rel_l2(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.double(), ref.double()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-300))


def max_abs_over_max_ref(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.double(), ref.double()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-300))
