"""Time gava_attention_bwd at the config-2 shape: `GAVA_LIB_PATH=... python tools/attn_bwd_bench.py`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gava_clip_b200 import ops
dev = torch.device("cuda:0")
B, T, N0, G, H = int(os.environ.get("B", 64)), int(os.environ.get("T", 8)), 197, 8, 12
D, BT = H * 64, B * T
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(BT * N0, 3 * D, device=dev, generator=g).bfloat16()
kv_aux = torch.randn(G + 2 * BT, 2 * D, device=dev, generator=g).bfloat16()
out = torch.empty(BT * N0, D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(BT, H, N0, device=dev)
ops.attention_train(qkv, kv_aux, out, lse, B=B, T=T, N0=N0, G=G, H=H)
dout = torch.randn(BT * N0, D, device=dev, generator=g).bfloat16()
dqkv = torch.empty_like(qkv); dkv = torch.empty(G + 2 * BT, 2 * D, device=dev)
slots = torch.empty(BT * H * (G + T + 1) * 128, device=dev)
for use_lse, use_slots in ((True, True), (True, False), (False, False)):
    kw = dict(B=B, T=T, N0=N0, G=G, H=H, lse=lse if use_lse else None, scratch=slots if use_slots else None)
    for _ in range(3):
        ops.attention_bwd(qkv, kv_aux, out, dout, dqkv, dkv, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention_bwd(qkv, kv_aux, out, dout, dqkv, dkv, **kw)
    e1.record(); torch.cuda.synchronize()
    print(f"{os.environ.get('GAVA_LIB_PATH', 'default')} T={T} lse={use_lse} slots={use_slots}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
