"""attention kernel alone on the config-2 shape (for ncu / timing).  usage: python tools/attn_one.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gava_clip_b200 import ops
B, T, N0, G, H = 64, 8, 197, 8, 12
D = H * 64
dev = torch.device("cuda")
qkv = torch.randn(B * T * N0, 3 * D, device=dev).bfloat16()
kv = torch.randn(G + 2 * B * T, 2 * D, device=dev).bfloat16()
out = torch.zeros(B * T * N0, D, device=dev, dtype=torch.bfloat16)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(2): ops.attention(qkv, kv, out, B=B, T=T, N0=N0, G=G, H=H)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): ops.attention(qkv, kv, out, B=B, T=T, N0=N0, G=G, H=H)
e1.record(); torch.cuda.synchronize()
print(f"attention: {e0.elapsed_time(e1)/n*1e3:.1f} us per launch (GAVA_ATTN_TC={os.environ.get('GAVA_ATTN_TC','1')})")
