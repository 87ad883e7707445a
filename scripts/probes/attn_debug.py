"""Attention kernel against a plain torch reference on a few geometries; prints where the two differ (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from llama32_b200 import ops
dev, dt = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g).to(dt)
def ref(q, k, v, kv_len, past, causal, keep):
    b, t, hd = q.shape
    kvh, d = k.shape[1], k.shape[3]
    heads = hd // d
    q4 = q.view(b, t, heads, d).transpose(1, 2).float()
    k4 = k[:, :, :kv_len].float().repeat_interleave(heads // kvh, 1)
    v4 = v[:, :, :kv_len].float().repeat_interleave(heads // kvh, 1)
    s = q4 @ k4.transpose(-1, -2) / d ** 0.5
    qi = torch.arange(t, device=dev)[:, None] + past
    kj = torch.arange(kv_len, device=dev)[None]
    m = torch.ones(t, kv_len, dtype=torch.bool, device=dev)
    if causal: m &= kj <= qi
    m = m[None, None].expand(b, 1, t, kv_len)
    if keep is not None: m = m & keep[:, None, None, :].bool()
    s = s.masked_fill(~m, float("-inf"))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p, nan=0.0)
    return (p @ v4).transpose(1, 2).reshape(b, t, hd)
for (b, t, heads, kvh, d, causal, pad) in [(1, 128, 2, 1, 128, True, 0), (1, 64, 2, 1, 128, True, 0), (1, 300, 4, 2, 128, True, 0),
                                           (2, 300, 4, 2, 128, True, 37), (2, 300, 4, 2, 64, True, 37), (2, 512, 8, 2, 128, False, 0)]:
    q, k, v = rnd(b, t, heads * d), rnd(b, kvh, t + 8, d), rnd(b, kvh, t + 8, d)
    keep = None
    if pad:
        keep = torch.ones(b, t, dtype=torch.uint8, device=dev); keep[0, :pad] = 0
    y = ops.gqa_attention_forward(q, k, v, t, 0, causal=causal, key_keep=keep).float()
    r = ref(q, k, v, t, 0, causal, keep)
    bad = ~torch.isfinite(y)
    err = (y - r).abs()
    err[bad] = 1e9
    rows = (err.view(b, t, heads, d).amax(-1) > 0.05).nonzero()
    print(f"b={b} t={t} heads={heads}/{kvh} d={d} causal={causal} pad={pad}: nonfinite {int(bad.sum())}, bad (b,row,head) count {rows.shape[0]}",
          rows[:6].tolist(), "max err", float(err[~bad].max()) if (~bad).any() else None, flush=True)
# detail of one padded case
b, t, heads, kvh, d = 1, 128, 1, 1, 128
for pad in (0, 5, 37, 64, 70):
    q, k, v = rnd(b, t, heads * d), rnd(b, kvh, t + 8, d), rnd(b, kvh, t + 8, d)
    keep = torch.ones(b, t, dtype=torch.uint8, device=dev); keep[0, :pad] = 0
    y = ops.gqa_attention_forward(q, k, v, t, 0, causal=True, key_keep=keep).float()
    r = ref(q, k, v, t, 0, True, keep)
    err = (y - r).abs().amax(-1)[0]
    badrows = (~(err < 0.05)).nonzero().flatten().tolist()
    print(f"pad={pad}: bad rows {badrows[:10]}... n={len(badrows)}")
    if badrows:
        i = badrows[0]
        print("  y", y[0, i, :6].tolist(), "\n  r", r[0, i, :6].tolist())
