"""Small fixed workloads for ncu: python scripts/profile_targets.py {decode|norm|prefill|train} (run plain first, then under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from llama32_b200 import ops  # noqa: E402

dt = torch.bfloat16
what = sys.argv[1]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
uni = lambda r, c: ((torch.rand(r, c, device=dev, generator=g) * 2 - 1) / c ** 0.5).to(dt)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g).to(dt)
if what == "decode":
    H, I = 4096, 14336
    ws = [(uni(I, H), uni(I, H), uni(H, I)) for _ in range(3)]
    gamma = torch.ones(H, device=dev, dtype=dt)
    for B in (1, 64):
        x, r = rnd(B, 1, H), rnd(B, 1, H)
        for i in range(6):
            wg, wu, wd = ws[i % 3]
            ops.ffn_forward(ops.add_rmsnorm_forward(x, gamma, r, 1e-5, want_rms=False)[0], wg, wu, wd)
elif what == "norm":
    T, H = 8192, 4096
    xs = [rnd(T, H) for _ in range(4)]
    rs = [rnd(T, H) for _ in range(4)]
    gamma = torch.ones(H, device=dev, dtype=dt)
    rms = torch.rand(T, device=dev, generator=g) + 0.5
    for i in range(6):
        ops.add_rmsnorm_forward(xs[i % 4], gamma, rs[(i + 1) % 4], 1e-5, want_rms=False)
        ops.add_rmsnorm_forward(xs[i % 4], gamma, rs[(i + 1) % 4], 1e-5, want_h=True)
        ops.rmsnorm_backward(xs[(i + 2) % 4], rs[(i + 3) % 4], gamma, rms)
elif what == "prefill":
    T, H, I = 8192, 4096, 14336
    wg, wu, wd = uni(I, H), uni(I, H), uni(H, I)
    x = rnd(T, H)
    for i in range(4):
        ops.ffn_forward(x, wg, wu, wd)
elif what == "train":
    T, H, I = 8192, 4096, 14336
    wg, wu, wd = uni(I, H), uni(I, H), uni(H, I)
    x, dy = rnd(T, H), rnd(T, H)
    for i in range(3):
        y, gc, uc = ops.ffn_forward(x, wg, wu, wd, want_cache=True)
        ops.ffn_backward(dy, x, wg, wu, wd, gc, uc)
elif what == "attention":
    B, T, NH, NKV, D = 4, 2048, 32, 8, 128
    q = rnd(B, T, NH * D)
    ck, cv = rnd(B, NKV, T, D), rnd(B, NKV, T, D)
    for i in range(3):
        ops.gqa_attention_forward(q, ck, cv, T, 0, causal=True)
    qd = rnd(64, 1, NH * D)
    ckd, cvd = rnd(64, NKV, 2112, D), rnd(64, NKV, 2112, D)
    for i in range(3):
        ops.gqa_attention_forward(qd, ckd, cvd, 2048, 2047, causal=True)
    kn, vn = rnd(B, T, NKV * D), rnd(B, T, NKV * D)
    pos = torch.arange(T, device=dev)[None].expand(B, -1).contiguous()
    for i in range(3):
        ops.rope_kv_append(q, kn, vn, pos, ck, cv, 0)
elif what == "lmhead":
    T, H, V = 8192, 4096, 128256
    w = uni(V, H)
    hs = rnd(T, H)
    labels = torch.randint(0, V, (T,), device=dev)
    for i in range(3):
        ops.lm_head_ce_forward(hs, w, labels)
torch.cuda.synchronize()
print("done", what)
