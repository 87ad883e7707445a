"""One decoder layer at the 11B geometry (norm1 -> GQA attention with RoPE + KV cache -> fused block tail): a prefill of 4 x 2048
tokens and one decode step of 64 sequences with 2048 cached tokens.  Run under `ncu --metrics gpu__time_duration.sum` for the
per-kernel times: python scripts/probes/layer_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import llama32_b200 as L
dev, dt = "cuda", torch.bfloat16
H, NH, NKV, I, EPS = 4096, 32, 8, 14336, 1e-5
class C: hidden_size, n_heads, n_kv_groups, rope_base = H, NH, NKV, 500000.0
class Layer(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.att = L.GroupQueryAttention(C, layer_idx=0)
        self.norm1, self.norm2 = L.LLAMARMSNorm(H, EPS), L.LLAMARMSNorm(H, EPS)
        self.ff = L.FusedFeedforward(H, I)
    def forward(self, hs, mask, pos, cache):
        a = self.att(self.norm1(hs), attention_mask=mask, position_ids=pos, kv_cache=cache)
        return L.block_tail(self.norm2, self.ff, a, hs)
layer = Layer().to(dev, dt).eval()
B, T = 4, 2048
hs = torch.randn(B, T, H, device=dev).to(dt)
pos = torch.arange(T, device=dev)[None].expand(B, -1).contiguous()
mask = torch.triu(torch.full((T, T), float("-inf"), device=dev, dtype=dt), diagonal=1)[None, None].expand(B, 1, T, T)
with torch.no_grad():
    for _ in range(2):
        layer(hs, mask, pos, L.KVCache(capacity=T))
    torch.cuda.synchronize()
    print("== prefill", flush=True)
    layer(hs, mask, pos, L.KVCache(capacity=T))
    torch.cuda.synchronize()
    Bd, Lk = 64, 2048
    cache = L.KVCache(capacity=Lk + 64)
    for i in range(0, Lk, 512):
        chunk = torch.randn(Bd, 512, H, device=dev).to(dt)
        cpos = torch.arange(i, i + 512, device=dev)[None].expand(Bd, -1).contiguous()
        layer.att(layer.norm1(chunk), attention_mask=None, position_ids=cpos, kv_cache=cache)
    x1 = torch.randn(Bd, 1, H, device=dev).to(dt)
    p1 = torch.full((Bd, 1), Lk, device=dev, dtype=torch.long)
    zmask = torch.zeros(Bd, 1, 1, 1, device=dev, dtype=dt)
    torch.cuda.synchronize()
    print("== decode", flush=True)
    layer(x1, zmask, p1, cache)
    torch.cuda.synchronize()
print("done")
