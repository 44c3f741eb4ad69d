"""attention kernel alone (for ncu / timing).  usage: [B=64 T=8 N0=197 H=12] python tools/attn_one.py [iters]
Defaults = config 2 (ViT-B/16, 8 frames, batch 64); ViT-L/14 T=16 batch 16: B=16 T=16 N0=257 H=16."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gava_clip_b200 import ops
B, T, N0, H = (int(os.environ.get(k, d)) for k, d in (("B", 64), ("T", 8), ("N0", 197), ("H", 12)))
G = 8
D = H * 64
dev = torch.device("cuda")
qkv = torch.randn(B * T * N0, 3 * D, device=dev).bfloat16()
kv = torch.randn(G + 2 * B * T, 2 * D, device=dev).bfloat16()
out = torch.zeros(B * T * N0, D, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(2): ops.attention(qkv, kv, out, B=B, T=T, N0=N0, G=G, H=H)
torch.cuda.synchronize()
ts = []
for _ in range(n):
    flush.zero_()                      # inputs (0.9 GB at config 2) exceed L2 anyway; keep launches independent
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.attention(qkv, kv, out, B=B, T=T, N0=N0, G=G, H=H); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
byt = (B * T * N0 * 4 * D + (G + 2 * B * T) * 2 * D) * 2
print(f"attention B={B} T={T} N0={N0} H={H}: median {ts[len(ts)//2]:.1f} us, min {ts[0]:.1f} us per launch; "
      f"{byt/1e6:.0f} MB algorithmic -> {byt/ts[len(ts)//2]/1e3:.0f} GB/s (GAVA_ATTN_TC={os.environ.get('GAVA_ATTN_TC','1')})")
