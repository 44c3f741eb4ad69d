import os, sys, torch
sys.path.insert(0, "/root/repo")
from gava_clip_b200 import ops
dev = torch.device("cuda:0")
B, T, N0, G, H = 64, 8, 197, 8, 12
D, BT = H * 64, B * T
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(BT * N0, 3 * D, device=dev, generator=g).bfloat16()
kv_aux = torch.randn(G + 2 * BT, 2 * D, device=dev, generator=g).bfloat16()
out = torch.zeros(BT * N0, D, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, kv_aux, out, B=B, T=T, N0=N0, G=G, H=H, n_q=1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attention(qkv, kv_aux, out, B=B, T=T, N0=N0, G=G, H=H, n_q=1)
e1.record(); torch.cuda.synchronize()
print("GAVA_ATTN_Q1=", os.environ.get("GAVA_ATTN_Q1", "1"), f"{e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
