"""CPU, build container only: run the REFERENCE's tiny MLLAMA (config 1) with its own modules and with ours
swapped in (patch_reference / convert_instances) -- identical logits; and re-derive the committed digest."""
import sys

import pytest
import torch

from conftest import REFERENCE, have_reference, load_golden

pytestmark = pytest.mark.skipif(not have_reference(), reason="reference tree not present (GPU box)")


def _build(seed=2024):
    sys.dont_write_bytecode = True
    for p in (REFERENCE, REFERENCE + "/Model"):
        if p not in sys.path:
            sys.path.insert(0, p)
    import Model.model as M
    torch.manual_seed(seed)
    vision_cfg = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, image_size=28,
                      patch_size=14)
    text_cfg = dict(vocab_size=512, hidden_size=256, n_heads=8, n_layers=2, hidden_dim=688, n_kv_groups=2,
                    dtype=torch.float32)
    cfg = M.MLLAMAConfig(vision_config=vision_cfg, text_config=text_cfg, projection_dim=256, image_token_index=511)
    model = M.MllamaForConditionalGeneration(cfg).eval()
    ids = torch.randint(0, 500, (2, 16))
    ids[:, :4] = 511
    pix = torch.randn(2, 3, 28, 28)
    return M, model, ids, pix


def _logits(model, ids, pix):
    with torch.no_grad():
        out = model(input_ids=ids, pixel_values=pix, attention_mask=torch.ones_like(ids))
    return out["logits"] if isinstance(out, dict) else out[0]


def test_config1_digest_and_drop_in():
    import llama32_b200 as L
    g = load_golden("mllama_cfg1.npz")
    M, model, ids, pix = _build()
    ref = _logits(model, ids, pix)
    assert list(ref.shape) == g["shape"].tolist()
    assert torch.equal(ref[0, :4, :8], g["logits_first"])
    assert abs(float(ref.double().sum()) - g["logits_sum"]) < 1e-6 * max(1.0, abs(g["logits_sum"]))
    keys_before = list(model.state_dict().keys())
    L.convert_instances(model)             # re-class the live instances: our modules now run the hot path
    n_ours = sum(isinstance(m, (L.LLAMARMSNorm, L.FusedFeedforward)) for m in model.modules())
    assert n_ours >= 2 * 2 + 1 + 2          # 2 norms + 1 ffn per layer, + final norm
    ours = _logits(model, ids, pix)
    assert torch.equal(ours, ref)
    assert list(model.state_dict().keys()) == keys_before
    # and training now works end to end on CPU, which the reference could not do (SwiGLUFunction.backward raised)
    out = model(input_ids=ids, pixel_values=pix, attention_mask=torch.ones_like(ids), labels=ids)
    loss = out["loss"] if isinstance(out, dict) else out[1]
    loss.backward()
    assert model.language_model.model.trf_blocks[0].ff.swiglu.w_gate.grad is not None


def test_patch_reference_builds_our_classes():
    import llama32_b200 as L
    M, _, _, _ = _build()
    import Tools.swiglu.FusedSwiglu as FS
    saved = (M.LLAMARMSNorm, M.FusedFeedforward, M.FusedSwiGLU, M.Linear_LORA, M.RMSNormFunction, M.HAS_RMSNORM_EXT,
             FS.SwiGLUFunction, FS.FusedSwiGLU, FS.FusedFeedForward, FS.CUDA_AVAILABLE)
    saved2 = (M.GroupQueryAttention, M.KVCache, M.TransformerBlock.forward, M.MllamaForConditionalGeneration.forward)
    try:
        L.patch_reference(M, FS)
        cfg = M.LLAMA32Config(vocab_size=64, hidden_size=32, n_heads=4, n_layers=1, hidden_dim=88, n_kv_groups=2,
                              dtype=torch.float32)
        blk = M.TransformerBlock(cfg, 0)
        assert isinstance(blk.norm2, L.LLAMARMSNorm) and isinstance(blk.ff, L.FusedFeedforward)
        assert isinstance(blk.att, L.GroupQueryAttention) and M.KVCache is L.KVCache
    finally:
        (M.GroupQueryAttention, M.KVCache, M.TransformerBlock.forward, M.MllamaForConditionalGeneration.forward) = saved2
        (M.LLAMARMSNorm, M.FusedFeedforward, M.FusedSwiGLU, M.Linear_LORA, M.RMSNormFunction, M.HAS_RMSNORM_EXT,
         FS.SwiGLUFunction, FS.FusedSwiGLU, FS.FusedFeedForward, FS.CUDA_AVAILABLE) = saved


def test_kv_cached_decode_and_loss_drop_in():
    """The reference's own model, prefill + one KV-cached decode step (explicit position_ids) + the loss with labels, against
    the same model with our classes swapped in and OUR preallocated KVCache: identical logits, identical loss (CPU fp32: our
    modules evaluate the reference's expressions there, so this pins the restated attention / cache / loss paths bit for bit)."""
    import llama32_b200 as L
    M, model, ids, _ = _build(seed=7)
    labels = ids.clone()
    labels[:, :5] = model.ignore_index
    nxt = torch.randint(0, 500, (2, 1))
    pos = torch.full((2, 1), ids.shape[1], dtype=torch.long)

    def run(cache):
        with torch.no_grad():
            a = model(input_ids=ids, attention_mask=torch.ones_like(ids), labels=labels, kv_cache=cache)
            b = model(input_ids=nxt, position_ids=pos, kv_cache=cache)
        return a["logits"], a["loss"], b["logits"]

    ref = run(M.KVCache())
    L.convert_instances(model)
    assert any(isinstance(m, L.GroupQueryAttention) for m in model.modules())
    ours_cache = L.KVCache()
    # the converted instances keep the reference's forward of the outer model: patch the loss tail by hand for this instance
    import types
    model.forward = types.MethodType(L.modules._mllama_forward, model)
    ours = run(ours_cache)
    assert ours_cache.num_items() == ids.shape[1] + 1
    for a, b, what in zip(ref, ours, ("prefill logits", "loss", "decode logits")):
        assert torch.equal(a, b), what
