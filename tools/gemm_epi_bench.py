"""Epilogue cost of the tcgen05 GEMM on the fine-tune / fold shapes (CUDA events, operands beyond L2):
   python tools/gemm_epi_bench.py          -> us and TFLOP/s per (shape, epilogue mode)
Compare with a GAVA_GEMM_DEBUG=1 variant (epilogue conversion + stores skipped) to separate mainloop from epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gava_clip_b200 import ops

dev = torch.device("cuda")
M = 100864
g = torch.Generator(device="cuda").manual_seed(0)


def run(name, N, K, mode, **extra):
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    f32 = mode in (ops.EPI_BIAS_RESID_F32, ops.EPI_BIAS_RESID_STATS_F32)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    kw = dict(mode=mode, bias=bias)
    if f32:
        kw["resid"] = out
        kw["aux"] = torch.randn(M, N, device=dev, generator=g).bfloat16()
    if mode == ops.EPI_BIAS_RESID_STATS_F32:
        kw["copy_out"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw["stats_out"] = torch.empty(M, ops.gemm_stats_slots(M, N), 2, device=dev)
    if mode == ops.EPI_DGELU_BF16:
        kw["aux"] = torch.randn(M, N, device=dev, generator=g).bfloat16()
    if mode == ops.EPI_BIAS_GELU_BF16 and extra.get("aux_out"):
        kw["aux_out"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if mode == ops.EPI_LNFOLD_BF16:
        kw["stats_in"] = torch.rand(M, K // 128, 2, device=dev, generator=g) + 1.0
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
    print(f"{name:34s} {ms * 1e3:7.1f} us {2.0 * M * N * K / ms / 1e9:7.1f} TF/s", flush=True)


print("lib:", os.environ.get("GAVA_LIB_PATH", "default"))
run("N3072 K768  bias->bf16 (mode 0)", 3072, 768, ops.EPI_BIAS_BF16)
run("N3072 K768  gelu (mode 1)", 3072, 768, ops.EPI_BIAS_GELU_BF16)
run("N3072 K768  gelu + aux_out", 3072, 768, ops.EPI_BIAS_GELU_BF16, aux_out=True)
run("N3072 K768  dgelu (mode 5)", 3072, 768, ops.EPI_DGELU_BF16)
run("N768  K3072 resid+delta (mode 2)", 768, 3072, ops.EPI_BIAS_RESID_F32)
run("N768  K3072 resid+stats (mode 8)", 768, 3072, ops.EPI_BIAS_RESID_STATS_F32)
run("N2304 K768  bias->bf16 (mode 0)", 2304, 768, ops.EPI_BIAS_BF16)
run("N2304 K768  ln-fold (mode 7)", 2304, 768, ops.EPI_LNFOLD_BF16)
run("N768  K768  bias->bf16 (mode 0)", 768, 768, ops.EPI_BIAS_BF16)
run("N768  K768  resid (mode 2)", 768, 768, ops.EPI_BIAS_RESID_F32)
