"""one GEMM shape, few launches (for ncu).  usage: python tools/gemm_one.py qkv|out|fc1|fc2"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gava_clip_b200 import ops
S = {"qkv": (100864, 2304, 768, ops.EPI_BIAS_BF16), "out": (100864, 768, 768, ops.EPI_BIAS_RESID_F32),
     "fc1": (100864, 3072, 768, ops.EPI_BIAS_GELU_BF16), "fc2": (100864, 768, 3072, ops.EPI_BIAS_RESID_F32)}
M, N, K, mode = S[sys.argv[1]]
dev = torch.device("cuda")
a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
bias = torch.randn(N, device=dev)
out = torch.zeros(M, N, device=dev, dtype=torch.float32 if mode == ops.EPI_BIAS_RESID_F32 else torch.bfloat16)
kw = dict(mode=mode, bias=bias)
if mode == ops.EPI_BIAS_RESID_F32: kw["resid"] = out
for _ in range(4): ops.gemm(a, w, out, **kw)
torch.cuda.synchronize(); print("ok")
