"""Attention kernel timing: prefill (11B geometry) and decode.  python scripts/attn_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llama32_b200 import ops
dev, dt = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g).to(dt)
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for B, T, NH, NKV, D in ((4, 2048, 32, 8, 128), (1, 8192, 32, 8, 128), (4, 2048, 32, 8, 64)):
    q, ck, cv = rnd(B, T, NH * D), rnd(B, NKV, T, D), rnd(B, NKV, T, D)
    ms = timeit(lambda: ops.gqa_attention_forward(q, ck, cv, T, 0, causal=True))
    fl = 4.0 * B * NH * T * T * D / 2
    q4 = q.view(B, T, NH, D).transpose(1, 2)
    ms_sdpa = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4, ck, cv, is_causal=True, enable_gqa=True))
    print(f"prefill B={B} T={T} heads {NH}/{NKV} d={D}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s   (torch SDPA {ms_sdpa:.3f} ms)", flush=True)
for B, Lk in ((64, 2048), (8, 8192), (1, 2048)):
    NH, NKV, D = 32, 8, 128
    qd, ck, cv = rnd(B, 1, NH * D), rnd(B, NKV, Lk + 64, D), rnd(B, NKV, Lk + 64, D)
    ms = timeit(lambda: ops.gqa_attention_forward(qd, ck, cv, Lk, Lk - 1, causal=True), iters=50)
    by = 2.0 * B * NKV * Lk * D * 2
    print(f"decode B={B} kv={Lk}: {ms * 1e3:.1f} us  {by / ms / 1e6:.0f} GB/s of K+V", flush=True)
# RoPE + cache append (11B geometry): prefill 4 x 2048 tokens and a decode step of 64 sequences
for B, T in ((4, 2048), (64, 1)):
    NH, NKV, D = 32, 8, 128
    qr, kn, vn = rnd(B, T, NH * D), rnd(B, T, NKV * D), rnd(B, T, NKV * D)
    ck, cv = torch.zeros(B, NKV, T + 64, D, device=dev, dtype=dt), torch.zeros(B, NKV, T + 64, D, device=dev, dtype=dt)
    pos = torch.arange(T, device=dev)[None].expand(B, -1).contiguous()
    ms = timeit(lambda: ops.rope_kv_append(qr, kn, vn, pos, ck, cv, 0), iters=50)
    by = (2 * B * T * NH * D + 4 * B * T * NKV * D) * 2
    print(f"rope + cache append B={B} T={T}: {ms * 1e3:.1f} us  {by / ms / 1e6:.0f} GB/s", flush=True)
# the same prefill with a key-padding vector (what a 4-D additive mask from the model becomes): no padded key, 100 padded keys
B, T, NH, NKV, D = 4, 2048, 32, 8, 128
q, ck, cv = rnd(B, T, NH * D), rnd(B, NKV, T, D), rnd(B, NKV, T, D)
for pad in (0, 100):
    keep = torch.ones(B, T, dtype=torch.uint8, device=dev); keep[:, :pad] = 0
    ms = timeit(lambda: ops.gqa_attention_forward(q, ck, cv, T, 0, causal=True, key_keep=keep))
    print(f"prefill B={B} T={T} with key padding ({pad} padded keys): {ms:.3f} ms", flush=True)
