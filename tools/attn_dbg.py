"""debug: attention_train on the test_attention_bwd data; which rows are wrong?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gava_clip_b200 import ops
from tests.test_backward_kernels_gpu import _attn_ref, _gen
cuda = torch.device("cuda")
B, T, N0, G, H = 2, 8, 197, 8, 12
scale = float(os.environ.get("SCALE", "1.5"))
g = _gen(B * 1000 + T * 10 + N0)
D, BT = H * 64, B * T
qkv = (torch.randn(BT * N0, 3 * D, device=cuda, generator=g) * scale).bfloat16()
kv_aux = (torch.randn(G + 2 * BT, 2 * D, device=cuda, generator=g) * scale).bfloat16()
ref = _attn_ref(qkv.float(), kv_aux.float(), B, T, N0, G, H)
for trial in range(3):
    for with_lse in (False, True):
        out = torch.zeros(BT * N0, D, device=cuda, dtype=torch.bfloat16)
        lse = torch.zeros(BT, H, N0, device=cuda)
        if with_lse: ops.attention_train(qkv, kv_aux, out, lse, B=B, T=T, N0=N0, G=G, H=H)
        else: ops.attention(qkv, kv_aux, out, B=B, T=T, N0=N0, G=G, H=H)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().view(BT, N0, H, 64).amax(-1)       # per (frame, row, head)
        bad = (err > 0.05) | ~torch.isfinite(err)
        print(f"trial {trial} lse={with_lse}: bad (frame,row,head) cells {int(bad.sum())} of {bad.numel()}; "
              f"bad per head {bad.sum((0,1)).tolist()}; bad rows<128: {int(bad[:, :128].sum())} rows>=128: {int(bad[:, 128:].sum())}")
        if bad.any():
            idx = bad.nonzero()[:6].tolist()
            print("  first bad cells", idx, "values", [float(out.float().view(BT, N0, H, 64)[i[0], i[1], i[2], 0]) for i in idx])
