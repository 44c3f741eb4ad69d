import sys, torch
sys.path.insert(0, "/root/repo")
from gava_clip_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
B, T, N0, G, H = 2, 8, 197, 8, 12
D, BT = H * 64, B * T
qkv = (torch.randn(BT * N0, 3 * D, device=dev, generator=g) * 1.5).bfloat16()
kv_aux = (torch.randn(G + 2 * BT, 2 * D, device=dev, generator=g) * 1.5).bfloat16()
out = torch.randn(BT * N0, D, device=dev, generator=g).bfloat16()
dout = torch.randn(BT * N0, D, device=dev, generator=g).bfloat16()
ref = None
bad = 0
for it in range(30):
    dqkv = torch.empty_like(qkv); dkv = torch.empty(G + 2 * BT, 2 * D, device=dev)
    ops.attention_bwd(qkv, kv_aux, out, dout, dqkv, dkv, B=B, T=T, N0=N0, G=G, H=H)
    torch.cuda.synchronize()
    if ref is None:
        ref = (dqkv.clone(), dkv.clone())
    else:
        if not torch.equal(dqkv, ref[0]):
            d = (dqkv.float() - ref[0].float()).abs()
            idx = d.nonzero()
            bad += 1
            print("attention_bwd dqkv differs: iter", it, "n", idx.shape[0], "max", float(d.max()), "first", idx[0].tolist(), "cols", sorted(set((idx[:, 1] // D).tolist())))
        dd = float((dkv - ref[1]).abs().max() / ref[1].abs().max())
        if dd > 1e-5:
            print("attention_bwd dkv_aux differs", it, dd)
print("attention_bwd mismatching iterations:", bad)
# forward tc attention determinism
o_ref = None
for it in range(20):
    o = torch.empty(BT * N0, D, device=dev, dtype=torch.bfloat16)
    ops.attention(qkv, kv_aux, o, B=B, T=T, N0=N0, G=G, H=H)
    torch.cuda.synchronize()
    if o_ref is None: o_ref = o.clone()
    elif not torch.equal(o, o_ref): print("attention fwd differs", it)
# GEMM variants
M = BT * N0
a = torch.randn(M, 768, device=dev, generator=g).bfloat16()
w = (torch.randn(3072, 768, device=dev, generator=g) * 0.03).bfloat16()
aux = torch.randn(M, 3072, device=dev, generator=g).bfloat16()
r = None
for it in range(20):
    o = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, o, mode=ops.EPI_DGELU_BF16, aux=aux)
    torch.cuda.synchronize()
    if r is None: r = o.clone()
    elif not torch.equal(o, r): print("gemm dgelu differs", it)
x = torch.randn(M, 768, device=dev, generator=g)
wn = torch.randn(768, device=dev, generator=g)
dy = torch.randn(M, 768, device=dev, generator=g).bfloat16()
r = None
for it in range(20):
    dx = torch.ones(M, 768, device=dev); dxb = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    ops.layernorm_bwd(dy, x, wn, dx, rows=M, D=768, add1=dx, dx_bf16=dxb)
    torch.cuda.synchronize()
    if r is None: r = dx.clone()
    elif not torch.equal(dx, r): print("ln bwd differs", it)
# small in-place strided GEMM (cls rows)
dcp = torch.randn(BT, 768, device=dev, generator=g).bfloat16()
wc = (torch.randn(768, 768, device=dev, generator=g) * 0.03).bfloat16()
base = torch.randn(M, 768, device=dev, generator=g)
r = None
for it in range(20):
    dx = base.clone()
    v = dx.view(BT, N0 * 768)[:, :768]
    ops.gemm(dcp, wc, v, mode=ops.EPI_BIAS_RESID_F32, resid=v)
    torch.cuda.synchronize()
    if r is None: r = dx.clone()
    elif not torch.equal(dx, r): print("cls gemm differs", it)
# temporal attention bwd
sq = torch.randn(BT, 3 * D, device=dev, generator=g).bfloat16()
do = torch.randn(BT, D, device=dev, generator=g).bfloat16()
r = None
for it in range(20):
    o = torch.empty_like(sq)
    ops.temporal_attention_bwd(sq, do, o, B, T, H)
    torch.cuda.synchronize()
    if r is None: r = o.clone()
    elif not torch.equal(o, r): print("temporal bwd differs", it)
print("done")
