"""small-M GEMMs of the temporal-message chain (CUDA events, 200 launches each)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gava_clip_b200 import ops
dev = torch.device("cuda")
for name, M, N, K, mode in [("cls_proj", 512, 768, 768, ops.EPI_BIAS_F32), ("sqkv", 512, 2304, 768, ops.EPI_BIAS_BF16),
                            ("sout", 512, 768, 768, ops.EPI_BIAS_RESID_F32), ("kvaux", 1032, 1536, 768, ops.EPI_BIAS_BF16),
                            ("last fc1", 512, 3072, 768, ops.EPI_BIAS_GELU_BF16), ("last fc2", 512, 768, 3072, ops.EPI_BIAS_RESID_F32)]:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if mode in (ops.EPI_BIAS_BF16, ops.EPI_BIAS_GELU_BF16) else torch.float32)
    kw = dict(mode=mode, bias=bias)
    if mode == ops.EPI_BIAS_RESID_F32: kw["resid"] = out
    g = torch.cuda.CUDAGraph()
    for _ in range(3): ops.gemm(a, w, out, **kw)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(50): ops.gemm(a, w, out, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name:9s} M={M:5d} N={N:5d} K={K:5d}: {e0.elapsed_time(e1)/100*1e3:6.2f} us", end=" | ")
print(os.environ.get("GAVA_SMALL_M", ""))
