"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REFERENCE ITSELF.

Run in the build container (where /root/reference is mounted, read-only):
    python oracle/make_golden.py
It imports the reference's own modules (Model/model.py, Tools/swiglu/FusedSwiglu.py), feeds them seeded
bf16-representable inputs, and stores inputs + the reference's fp32 outputs.  The GPU box has no
/root/reference; tests read only the committed fixtures.

What the reference can and cannot produce:
  * forward of LLAMARMSNorm / FusedSwiGLU / FusedFeedforward / Linear_LORA: the live PyTorch branch -> stored;
  * gradients of LLAMARMSNorm: its fallback branch is plain autograd-differentiable torch -> stored;
  * gradients of the FFN: SwiGLUFunction.backward raises on every path (it unpacks tensors the fallback
    never saved), so no reference gradient exists; the fixture stores autograd over F.linear/F.silu written
    out exactly as FusedSwiglu.py:18-20 + model.py:217 and says so in its `grad_source` field.
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = os.environ.get("L32_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "Model"))
import Model.model as M  # noqa: E402
import Tools.swiglu.FusedSwiglu as FS  # noqa: E402
import torch.nn.functional as F  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def rep(t):   # bf16-representable fp32
    return t.to(torch.bfloat16).to(torch.float32)


def bits(t):  # bf16 bit pattern as uint16 (halves the fixture size; exact because t is bf16-representable)
    return t.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrs.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def golden_rmsnorm():
    torch.manual_seed(1234)
    eps = 1e-5
    for tag, shape in (("small", (3, 5, 64)), ("odd", (7, 250)), ("cfg1", (2, 16, 256))):
        c = shape[-1]
        x, r, g = rep(torch.randn(*shape)), rep(torch.randn(*shape)), rep(torch.randn(*shape))
        norm = M.LLAMARMSNorm(c, eps=eps)
        with torch.no_grad():
            norm.weight.copy_(rep(1 + 0.1 * torch.randn(c)))
        out = {}
        for res_tag, res in (("nores", None), ("res", r)):
            xs = x.clone().requires_grad_(True)
            rs = None if res is None else res.clone().requires_grad_(True)
            norm.weight.grad = None
            y = norm(xs, residual=rs)          # reference fallback branch, model.py:166-171
            y.backward(g)
            out[f"y_{res_tag}"] = y
            out[f"dx_{res_tag}"] = xs.grad
            out[f"dw_{res_tag}"] = norm.weight.grad.clone()
            if rs is not None:
                out["dres_res"] = rs.grad
        save(f"rmsnorm_{tag}.npz", x=x, residual=r, grad_out=g, weight=norm.weight.detach(), eps=np.float32(eps),
             source="reference Model/model.py:158-171 LLAMARMSNorm (fallback branch), grads by autograd through it",
             **out)


def golden_ffn():
    torch.manual_seed(4321)
    for tag, (shape, inter, bias) in (("small", ((2, 5, 64), 176, False)), ("bias", ((4, 64), 104, True)),
                                      ("cfg1", ((2, 16, 256), 688, False))):
        hidden = shape[-1]
        ff = M.FusedFeedforward(hidden, inter, bias=bias)
        with torch.no_grad():
            for p in ff.parameters():
                if p.dim() == 1:
                    p.copy_(rep(0.1 * torch.randn_like(p)))
                else:
                    p.copy_(rep(p))
        x, gy = rep(torch.randn(*shape)), rep(torch.randn(*shape))
        with torch.no_grad():
            act = ff.swiglu(x)              # SwiGLUFunction fallback branch, FusedSwiglu.py:17-20
            y = ff(x)                       # model.py:217
            # the class of the same name in Tools/swiglu (capital F) must agree
            ff2 = FS.FusedFeedForward(hidden, inter, bias=bias)
            ff2.load_state_dict(ff.state_dict())
            assert torch.equal(ff2(x), y)
        sd = ff.state_dict()
        wg, wu, wd = sd["swiglu.w_gate"], sd["swiglu.w_up"], sd["w_down.weight"]
        # gradients: no reference backward exists -> autograd over the same expressions
        xs, gs, us, ds = (t.clone().requires_grad_(True) for t in (x, wg, wu, wd))
        bg = sd.get("swiglu.b_gate"); bu = sd.get("swiglu.b_up"); bd = sd.get("w_down.bias")
        yy = F.linear(F.silu(F.linear(xs, gs, bg)) * F.linear(xs, us, bu), ds, bd)
        assert torch.equal(yy.detach(), y)
        yy.backward(gy)
        extra = {}
        if bias:
            extra = dict(b_gate=bg, b_up=bu, b_down=bd)
        # weight gradients are as large as the weights: keep them only for the small fixtures
        wgrads = {} if tag == "cfg1" else dict(dw_gate=gs.grad, dw_up=us.grad, dw_down=ds.grad)
        save(f"ffn_{tag}.npz", x=x, grad_out=gy, w_gate_bits=bits(wg), w_up_bits=bits(wu), w_down_bits=bits(wd),
             act=act, y=y, dx=xs.grad, **wgrads,
             source="reference Model/model.py:210-217 FusedFeedforward -> Tools/swiglu/FusedSwiglu.py:17-20",
             grad_source="autograd over F.linear/F.silu as written at FusedSwiglu.py:18-20 + model.py:217 "
                         "(the reference's SwiGLUFunction.backward raises on every path)", **extra)


def golden_block():
    """Hot-path slice of TransformerBlock.forward (model.py:265-273) at config-1 dims, weights of the real module."""
    torch.manual_seed(99)
    cfg = M.LLAMA32Config(vocab_size=512, hidden_size=256, n_heads=8, n_layers=2, hidden_dim=688, n_kv_groups=2,
                          dtype=torch.float32)
    blk = M.TransformerBlock(cfg, 0)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(rep(p))
        blk.norm2.weight.copy_(rep(1 + 0.1 * torch.randn(256)))
    hidden = rep(torch.randn(2, 16, 256))
    attn_out = rep(torch.randn(2, 16, 256))
    with torch.no_grad():
        normed = blk.norm2(attn_out, residual=hidden)      # model.py:271
        ff_out = blk.ff(normed)                            # model.py:272
        block_out = attn_out + ff_out                      # model.py:273
    sd = blk.state_dict()
    save("block_cfg1.npz", hidden=hidden, attn_out=attn_out, norm2_weight=sd["norm2.weight"],
         eps=np.float32(blk.norm2.eps), w_gate_bits=bits(sd["ff.swiglu.w_gate"]), w_up_bits=bits(sd["ff.swiglu.w_up"]),
         w_down_bits=bits(sd["ff.w_down.weight"]), normed=normed, ff_out=ff_out, block_out=block_out,
         state_dict_keys=np.array(sorted(sd.keys())),
         source="reference Model/model.py:257-273 TransformerBlock: norm2(attn_out, residual) -> ff -> attn_out + ff_out")


def golden_lora():
    torch.manual_seed(7)
    lin = M.Linear_LORA(176, 64, rank=16, alpha=32.0, dropout=0.0)
    with torch.no_grad():
        for p in lin.parameters():
            p.copy_(rep(p))
        lin.lora_b.weight.copy_(rep(0.05 * torch.randn_like(lin.lora_b.weight)))
    x = rep(torch.randn(6, 176))
    with torch.no_grad():
        y = lin(x)
    sd = lin.state_dict()
    save("lora_small.npz", x=x, y=y, w=sd["linear.weight"], lora_a=sd["lora_a.weight"], lora_b=sd["lora_b.weight"],
         alpha=np.float32(32.0), rank=np.int32(16), state_dict_keys=np.array(sorted(sd.keys())),
         source="reference Model/model.py:107-121 Linear_LORA (dropout p=0)")


def golden_mllama_cfg1():
    """Config 1: tiny random-init MLLAMA forward on CPU fp32. Stores only a digest of the
    logits plus the seed recipe; tests/test_reference_harness.py re-runs it when the reference is present."""
    torch.manual_seed(2024)
    vision_cfg = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, image_size=28,
                      patch_size=14)
    text_cfg = dict(vocab_size=512, hidden_size=256, n_heads=8, n_layers=2, hidden_dim=688, n_kv_groups=2,
                    dtype=torch.float32)
    cfg = M.MLLAMAConfig(vision_config=vision_cfg, text_config=text_cfg, projection_dim=256, image_token_index=511)
    model = M.MllamaForConditionalGeneration(cfg).eval()
    ids = torch.randint(0, 500, (2, 16))
    ids[:, :4] = 511
    pix = torch.randn(2, 3, 28, 28)
    with torch.no_grad():
        out = model(input_ids=ids, pixel_values=pix, attention_mask=torch.ones_like(ids))
    logits = out["logits"] if isinstance(out, dict) else out[0]
    save("mllama_cfg1.npz", input_ids=ids, logits_sum=np.float64(logits.double().sum()),
         logits_abs_sum=np.float64(logits.double().abs().sum()), logits_first=logits[0, :4, :8],
         shape=np.array(logits.shape), seed=np.int64(2024),
         source="reference Model/model.py:398-440 MllamaForConditionalGeneration.forward, config 1")


def golden_lm_head_loss():
    """The tail of the reference's own MllamaForConditionalGeneration.forward (Model/model.py:429-438): logits and the
    shifted cross-entropy loss of the config-1 model WITH labels (some of them ignore_index)."""
    torch.manual_seed(2025)
    vision_cfg = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, image_size=28,
                      patch_size=14)
    text_cfg = dict(vocab_size=512, hidden_size=256, n_heads=8, n_layers=2, hidden_dim=688, n_kv_groups=2,
                    dtype=torch.float32)
    cfg = M.MLLAMAConfig(vision_config=vision_cfg, text_config=text_cfg, projection_dim=256, image_token_index=511)
    model = M.MllamaForConditionalGeneration(cfg).eval()
    with torch.no_grad():
        model.language_model.lm_head.weight.copy_(rep(model.language_model.lm_head.weight))
    ids = torch.randint(0, 500, (3, 24))
    labels = ids.clone()
    labels[:, :5] = model.ignore_index
    labels[1, 17:] = model.ignore_index
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=torch.ones_like(ids), labels=labels)
    hs = rep(out["hidden_states"])
    # the same tail on bf16-representable hidden states (what the fixture stores), through the reference's own modules
    with torch.no_grad():
        logits = model.language_model.lm_head(hs)
        loss = torch.nn.CrossEntropyLoss(ignore_index=model.ignore_index)(
            logits[..., :-1, :].contiguous().view(-1, logits.size(-1)), labels[..., 1:].contiguous().view(-1))
    save("lm_head_ce_cfg1.npz", hidden_states=hs, lm_head_weight_bits=bits(model.language_model.lm_head.weight), labels=labels,
         ignore_index=np.int64(model.ignore_index), logits=logits, loss=np.float32(loss.item()),
         model_loss_unrounded_hidden=np.float32(out["loss"].item()),
         source="reference Model/model.py:429-438 (lm_head + shifted CrossEntropyLoss), config-1 MLLAMA, labels with ignore_index")


def golden_attention():
    """The reference's own GroupQueryAttention + KVCache + LLAMARotaryEmbedding + Llama3Model._prepare_attention_mask
    (Model/model.py:12-29, 176-198, 220-254, 304-319): a prefill with a padded batch, then one KV-cached decode step with
    EXPLICIT position_ids.  Two head geometries the tcgen05 kernel supports: head_dim 64 and 128."""
    for tag, hidden, heads, kv, t in (("d64", 256, 4, 2, 40), ("d128", 512, 4, 1, 150)):
        torch.manual_seed(77)
        cfg = M.LLAMA32Config(vocab_size=64, hidden_size=hidden, n_heads=heads, n_layers=1, hidden_dim=128, n_kv_groups=kv,
                              dtype=torch.float32)
        att = M.GroupQueryAttention(cfg, layer_idx=0, dtype=torch.float32).eval()
        lm = M.Llama3Model(cfg)                                   # only for its own mask / position helpers
        with torch.no_grad():
            for p_ in att.parameters():
                p_.copy_(rep(p_))
        b = 2
        x = rep(torch.randn(b, t, hidden))
        mask2d = torch.ones(b, t)
        mask2d[1, t - 7:] = 0                                     # right-padded second sequence
        mask4d = lm._prepare_attention_mask(mask2d, x)
        pos = lm._prepare_position_ids(None, x)
        cache = M.KVCache()
        with torch.no_grad():
            y = att(x, attention_mask=mask4d, position_ids=pos, kv_cache=cache)
            x1 = rep(torch.randn(b, 1, hidden))
            mask1 = lm._prepare_attention_mask(None, x1)          # what the reference builds for a decode step: [B,1,1,1] zeros
            pos1 = torch.full((b, 1), t, dtype=torch.long)        # explicit: the cache holds t tokens
            y1 = att(x1, attention_mask=mask1, position_ids=pos1, kv_cache=cache)
        sd = att.state_dict()
        save(f"attention_{tag}.npz", x=x, mask2d=mask2d, position_ids=pos, y_prefill=y, x_decode=x1, position_ids_decode=pos1,
             y_decode=y1, cache_k=cache.key_cache[0], cache_v=cache.value_cache[0],
             wq_bits=bits(sd["W_query.weight"]), wk_bits=bits(sd["W_key.weight"]), wv_bits=bits(sd["W_value.weight"]),
             wo_bits=bits(sd["out_proj.weight"]), n_heads=np.int32(heads), n_kv=np.int32(kv), rope_base=np.float32(cfg.rope_base),
             state_dict_keys=np.array(sorted(sd.keys())),
             source="reference Model/model.py:220-254 GroupQueryAttention with KVCache (:12-29), prefill + one decode step")


if __name__ == "__main__":
    golden_attention()
    golden_lm_head_loss()
    golden_rmsnorm()
    golden_ffn()
    golden_block()
    golden_lora()
    golden_mllama_cfg1()
