"""Micro-benchmark of the tcgen05 GEMM on the hot-path shapes (CUDA events, L2-exceeding operands).
   GAVA_LIB_PATH=<variant .so> python tools/gemm_bench.py   -> TFLOP/s per shape"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gava_clip_b200 import ops

SHAPES = [("qkv", 100864, 2304, 768, ops.EPI_BIAS_BF16), ("out", 100864, 768, 768, ops.EPI_BIAS_RESID_F32),
          ("fc1", 100864, 3072, 768, ops.EPI_BIAS_GELU_BF16), ("fc2", 100864, 768, 3072, ops.EPI_BIAS_RESID_F32)]
dev = torch.device("cuda")
res = []
for name, M, N, K, mode in SHAPES:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if mode == ops.EPI_BIAS_RESID_F32 else torch.bfloat16)
    kw = dict(mode=mode, bias=bias)
    if mode == ops.EPI_BIAS_RESID_F32:
        kw["resid"] = out
    for _ in range(3):
        ops.gemm(a, w, out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ops.gemm(a, w, out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res.append(f"{name}: {ms*1e3:7.1f} us {2.0*M*N*K/ms/1e9:7.1f} TF/s")
print(os.environ.get("GAVA_LIB_PATH", "default"), " | ".join(res))
