"""Per-kernel parity on the B200: each C-ABI entry point against a plain PyTorch fp32 reference
of the same op (the op as the reference model computes it; file:line in include/gava_sm100.h).

Tolerances: bf16 tensor-core inputs with fp32 accumulation -> compare against an fp32 reference fed
the SAME bf16-rounded operands, so only accumulation order / output rounding differ.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    return float((a.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 768, 768), (1000, 256, 128), (1032, 1536, 768),
                                   (4096, 2304, 768), (197 * 16, 768, 3072), (512, 128, 768), (333, 384, 192)])
def test_gemm_bias_bf16(cuda, M, N, K):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    w = (torch.randn(N, K, device=cuda, generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g)
    out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, w, out, mode=ops.EPI_BIAS_BF16, bias=bias)
    ref = a.float() @ w.float().t() + bias
    torch.cuda.synchronize()
    assert _rel(out.float(), ref) < 1e-2  # bf16 output rounding: 2^-8 relative


def test_gemm_gelu(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 640, 3072, 768
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    w = (torch.randn(N, K, device=cuda, generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g) * 0.1
    out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, w, out, mode=ops.EPI_BIAS_GELU_BF16, bias=bias)
    y = a.float() @ w.float().t() + bias
    ref = y * torch.sigmoid(1.702 * y)
    assert _rel(out.float(), ref) < 1e-2


def test_gemm_resid_inplace_and_strided_a(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    M, N, K = 1970, 768, 3072
    big = torch.randn(M, K + 64, device=cuda, generator=g).bfloat16()
    a = big[:, :K]                      # row pitch K+64: exercises lda != K
    w = (torch.randn(N, K, device=cuda, generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g)
    x = torch.randn(M, N, device=cuda, generator=g)
    ref = x + a.float() @ w.float().t() + bias
    ops.gemm(a, w, x, mode=ops.EPI_BIAS_RESID_F32, bias=bias, resid=x)
    assert _rel(x, ref) < 2e-5


def test_gemm_f32_no_bias(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    M, N, K = 512, 512, 768
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    w = (torch.randn(N, K, device=cuda, generator=g) * K ** -0.5).bfloat16()
    out = torch.empty(M, N, device=cuda)
    ops.gemm(a, w, out, mode=ops.EPI_BIAS_F32)
    assert _rel(out, a.float() @ w.float().t()) < 2e-5


def test_gemm_patch_epilogue(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    BT, T, Np, D, K = 6, 3, 196, 768, 768
    a = torch.randn(BT * Np, K, device=cuda, generator=g).bfloat16()
    w = (torch.randn(D, K, device=cuda, generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(D, device=cuda, generator=g)
    pos = torch.randn(Np + 1, D, device=cuda, generator=g)
    te = torch.randn(T, D, device=cuda, generator=g)
    x = torch.full((BT * (Np + 1), D), float("nan"), device=cuda)
    ops.gemm(a, w, x, mode=ops.EPI_PATCH_F32, bias=bias, pos=pos, time_embed=te, group=Np, group_stride=Np + 1,
             row_offset=1, T=T)
    y = (a.float() @ w.float().t() + bias).view(BT, Np, D) + pos[1:] + te[torch.arange(BT, device=cuda) % T][:, None]
    xv = x.view(BT, Np + 1, D)
    assert torch.isnan(xv[:, 0]).all()           # cls slots untouched
    assert _rel(xv[:, 1:], y) < 2e-5


def test_gemm_rejects_bad_shape(cuda):
    from gava_clip_b200 import ops
    a = torch.zeros(128, 100, device=cuda, dtype=torch.bfloat16)
    w = torch.zeros(128, 100, device=cuda, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.gemm(a, w, torch.empty(128, 128, device=cuda, dtype=torch.bfloat16), mode=ops.EPI_BIAS_BF16)


@pytest.mark.parametrize("D", [512, 768, 1024])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_layernorm(cuda, D, out_dtype):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(D)
    rows = 1001
    x = torch.randn(rows, D, device=cuda, generator=g) * 3 + 1
    w = torch.randn(D, device=cuda, generator=g)
    b = torch.randn(D, device=cuda, generator=g)
    out = torch.empty(rows, D, device=cuda, dtype=out_dtype)
    ops.layernorm(x, w, b, out, rows=rows, in_row_stride=D)
    ref = F.layer_norm(x, (D,), w, b, 1e-5)
    tol = 1e-2 if out_dtype == torch.bfloat16 else 1e-5
    assert _rel(out.float(), ref) < tol


def test_layernorm_gather_cast_and_inplace(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    BT, N0, D = 24, 197, 768
    x = torch.randn(BT * N0, D, device=cuda, generator=g)
    out = torch.empty(BT, D, device=cuda, dtype=torch.bfloat16)
    ops.layernorm(x, None, None, out, rows=BT, in_row_stride=N0 * D)       # gather cls rows, cast only
    assert torch.equal(out, x.view(BT, N0, D)[:, 0].bfloat16())
    w = torch.randn(D, device=cuda, generator=g); b = torch.randn(D, device=cuda, generator=g)
    ref = F.layer_norm(x, (D,), w, b, 1e-5)
    ops.layernorm(x, w, b, x, rows=BT * N0, in_row_stride=D)               # in place, fp32
    assert _rel(x, ref) < 1e-5


@pytest.mark.parametrize("P,H,Kpad", [(16, 224, 768), (14, 224, 640), (16, 64, 768)])
def test_patch_gather(cuda, P, H, Kpad):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(P)
    B, T = 2, 3
    v = torch.randn(B, 3, T, H, H, device=cuda, generator=g)
    gh = H // P
    out = torch.full((B * T * gh * gh, Kpad), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.patch_gather(v, out, P)
    frames = v.permute(0, 2, 1, 3, 4).flatten(0, 1)
    ref = F.unfold(frames, kernel_size=P, stride=P).transpose(1, 2).reshape(-1, 3 * P * P).bfloat16()
    assert torch.equal(out[:, :3 * P * P], ref)
    if Kpad > 3 * P * P:
        assert (out[:, 3 * P * P:] == 0).all()


def test_cls_rows(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    BT, T, N0, D = 12, 4, 197, 768
    x = torch.zeros(BT * N0, D, device=cuda)
    cls = torch.randn(D, device=cuda, generator=g); pos = torch.randn(N0, D, device=cuda, generator=g)
    te = torch.randn(T, D, device=cuda, generator=g)
    ops.cls_rows(x, N0, cls, pos, te, BT, T)
    ref = cls + pos[0] + te[torch.arange(BT, device=cuda) % T]
    assert torch.allclose(x.view(BT, N0, D)[:, 0], ref, atol=1e-6)
    assert (x.view(BT, N0, D)[:, 1:] == 0).all()


@pytest.mark.parametrize("T", [8, 16, 17, 70, 128])      # > 16 frames: the mma.sync kernel (seq_attention_mma.cu)
def test_temporal_attention(cuda, T):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(T)
    B, H = 3, 12
    D = H * 64
    qkv = torch.randn(B * T, 3 * D, device=cuda, generator=g).bfloat16()
    out = torch.empty(B * T, D, device=cuda, dtype=torch.bfloat16)
    ops.temporal_attention(qkv, out, B, T, H)
    q, k, v = qkv.float().view(B, T, 3, H, 64).unbind(2)
    aff = torch.einsum("nqhc,nkhc->nqkh", q / 8.0, k).softmax(dim=-2)
    ref = torch.einsum("nqlh,nlhc->nqhc", aff, v).reshape(B * T, D)
    assert _rel(out.float(), ref) < 1e-2


def test_prompt_rows(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    B, T, G, D = 3, 8, 8, 768
    gp = torch.randn(G, D, device=cuda, generator=g); lp = torch.randn(1, T, D, device=cuda, generator=g)
    cp = torch.randn(B * T, D, device=cuda, generator=g); sm = torch.randn(B * T, D, device=cuda, generator=g)
    w = torch.randn(D, device=cuda, generator=g); b = torch.randn(D, device=cuda, generator=g)
    out = torch.empty(G + 2 * B * T, D, device=cuda, dtype=torch.bfloat16)
    ops.prompt_rows(gp, lp, cp, sm, w, b, out, B, T)
    rows = torch.cat([gp, (lp.expand(B, -1, -1) + cp.view(B, T, D)).reshape(B * T, D), sm])
    ref = F.layer_norm(rows, (D,), w, b, 1e-5)
    assert _rel(out.float(), ref) < 1e-2


def _attention_ref(qkv, kv_aux, B, T, N0, G, H):
    D = H * 64
    BT = B * T
    q, k, v = qkv.float().view(BT, N0, 3, H, 64).unbind(2)
    ka, va = kv_aux.float().view(-1, 2, H, 64).unbind(1)
    outs = []
    for f in range(BT):
        b = f // T
        idx = list(range(G)) + list(range(G + b * T, G + (b + 1) * T)) + [G + BT + f]
        kk = torch.cat([k[f], ka[idx]]); vv = torch.cat([v[f], va[idx]])
        aff = torch.einsum("qhc,khc->qkh", q[f] / 8.0, kk).softmax(dim=1)
        outs.append(torch.einsum("qkh,khc->qhc", aff, vv).reshape(N0, D))
    return torch.stack(outs).reshape(BT * N0, D)


@pytest.mark.parametrize("T,N0,H", [(8, 197, 12), (16, 197, 12), (16, 257, 16), (3, 50, 2), (70, 197, 2), (24, 197, 4),
                                    (32, 257, 2), (5, 129, 3)])
def test_attention(cuda, T, N0, H):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(T * N0)
    B, G = 2, 8
    D = H * 64
    qkv = torch.randn(B * T * N0, 3 * D, device=cuda, generator=g).bfloat16()
    kv_aux = torch.randn(G + 2 * B * T, 2 * D, device=cuda, generator=g).bfloat16()
    out = torch.zeros(B * T * N0, D, device=cuda, dtype=torch.bfloat16)
    ops.attention(qkv, kv_aux, out, B=B, T=T, N0=N0, G=G, H=H)
    ref = _attention_ref(qkv, kv_aux, B, T, N0, G, H)
    assert _rel(out.float(), ref) < 1.5e-2   # P is rounded to bf16 before PV


def _lse_ref(qkv, kv_aux, B, T, N0, G, H):
    """log2-domain log-sum-exp of every query row, [BT, H, N0]"""
    D, BT = H * 64, B * T
    q, k, _ = qkv.float().view(BT, N0, 3, H, 64).unbind(2)
    ka = kv_aux.float().view(-1, 2, H, 64)[:, 0]
    rows = []
    for f in range(BT):
        b = f // T
        idx = list(range(G)) + list(range(G + b * T, G + (b + 1) * T)) + [G + BT + f]
        kk = torch.cat([k[f], ka[idx]])
        aff = torch.einsum("qhc,khc->hqk", q[f] / 8.0, kk)
        rows.append(torch.logsumexp(aff, dim=-1) * math.log2(math.e))
    return torch.stack(rows)


@pytest.mark.parametrize("T,N0,H", [(8, 197, 12), (16, 257, 4)])
def test_attention_single_pass_overflow_rows(cuda, T, N0, H):
    """The tcgen05 kernel shifts every row by the maximum of its FIRST 32 scores; rows where a later key beats that by
    more than 2^100 are recomputed exactly on the CUDA cores.  Craft such rows (a late key aligned with the query at a
    huge norm), moderately peaked rows (stale shift, no overflow) and ordinary rows; output and log-sum-exp must match."""
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(77)
    B, G = 2, 8
    D, BT = H * 64, B * T
    qkv = torch.randn(BT * N0, 3 * D, device=cuda, generator=g)
    v5 = qkv.view(BT, N0, 3, H, 64)
    # frame 1, head 0: query rows 3, 130 and (ViT-L) 256 see key 150 with logit 64*6*48 = 18432 (x0.18 = 3300 log2 units: overflow)
    hot_rows = [3, 130] + ([256] if N0 > 256 else [])
    for r in hot_rows:
        v5[1, r, 0, 0] = 6.0
    v5[1, 150, 1, 0] = 6.0 * 8.0
    # frame 2, head 1: key 180 beats the first chunk by ~40 log2 units for every row (stale shift, still exact)
    v5[2, :, 0, 1] = 1.0
    v5[2, 180, 1, 1] = 3.5
    # frame 0, last head: a PROMPT key (summary token) dominates rows 0..9
    kv_aux = torch.randn(G + 2 * BT, 2 * D, device=cuda, generator=g)
    v5[0, :10, 0, H - 1] = 5.0
    kv_aux.view(-1, 2, H, 64)[G + BT + 0, 0, H - 1] = 40.0
    qkv, kv_aux = qkv.bfloat16(), kv_aux.bfloat16()
    out = torch.zeros(BT * N0, D, device=cuda, dtype=torch.bfloat16)
    lse = torch.zeros(BT, H, N0, device=cuda)
    ops.attention_train(qkv, kv_aux, out, lse, B=B, T=T, N0=N0, G=G, H=H)
    ref = _attention_ref(qkv, kv_aux, B, T, N0, G, H)
    assert torch.isfinite(out.float()).all()
    assert _rel(out.float(), ref) < 1.5e-2
    lref = _lse_ref(qkv, kv_aux, B, T, N0, G, H)
    assert float(((lse - lref).abs() / lref.abs().clamp_min(1.0)).max()) < 2e-3
    for r in hot_rows:     # the recomputed rows individually
        a = out.float().view(BT, N0, H, 64)[1, r, 0]
        assert _rel(a, ref.view(BT, N0, H, 64)[1, r, 0]) < 1.5e-2


@pytest.mark.parametrize("B,T,N0,G,H", [(2, 8, 197, 8, 12), (1, 70, 197, 8, 2), (3, 3, 17, 8, 2), (1, 16, 257, 8, 16)])
def test_attention_cls_only(cuda, B, T, N0, G, H):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    D = H * 64
    qkv = torch.randn(B * T * N0, 3 * D, device=cuda, generator=g).bfloat16()
    kv_aux = torch.randn(G + 2 * B * T, 2 * D, device=cuda, generator=g).bfloat16()
    out = torch.zeros(B * T * N0, D, device=cuda, dtype=torch.bfloat16)
    ops.attention(qkv, kv_aux, out, B=B, T=T, N0=N0, G=G, H=H, n_q=1)
    ref = _attention_ref(qkv, kv_aux, B, T, N0, G, H).view(B * T, N0, D)
    o = out.float().view(B * T, N0, D)
    assert _rel(o[:, 0], ref[:, 0]) < 1.5e-2
    assert (o[:, 1:] == 0).all()
    # the single-query kernel also saves the log-sum-exp row for the backward; must agree with the full-tile kernels
    lse1 = torch.zeros(B * T, H, N0, device=cuda); lse_full = torch.zeros(B * T, H, N0, device=cuda)
    out2 = torch.zeros_like(out)
    ops.attention_train(qkv, kv_aux, out2, lse1, B=B, T=T, N0=N0, G=G, H=H, n_q=1)
    ops.attention_train(qkv, kv_aux, torch.zeros_like(out), lse_full, B=B, T=T, N0=N0, G=G, H=H)
    assert torch.equal(out2, out)
    assert float((lse1[:, :, 0] - lse_full[:, :, 0]).abs().max()) < 2e-3


@pytest.mark.parametrize("n_kv", [1, 5])
def test_head(cuda, n_kv):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(13)
    B, T, E, D, n_cls = 5, 8, 512, 768, 3
    ff = torch.randn(B * T, E, device=cuda, generator=g); st = torch.randn(B * T, D, device=cuda, generator=g)
    txt = torch.randn(n_cls * n_kv, E, device=cuda, generator=g)
    txt = txt / txt.norm(dim=-1, keepdim=True)
    offs = torch.arange(0, n_cls * n_kv + 1, n_kv, device=cuda, dtype=torch.int32)
    vf = torch.empty(B, E, device=cuda); so = torch.empty(B, D, device=cuda); lg = torch.empty(B, n_cls, device=cuda)
    scale = math.exp(math.log(1 / 0.07))
    ops.head(ff, st, txt, offs, vf, so, lg, B=B, T=T, scale=scale, logit_bias=-0.5)
    v = ff.view(B, T, E).mean(1)
    vh = v / v.norm(dim=-1, keepdim=True)
    ref = (scale * vh @ txt.t()).view(B, n_cls, n_kv).mean(-1) - 0.5
    assert _rel(vf, v) < 1e-6 and _rel(so, st.view(B, T, D).mean(1)) < 1e-6
    assert _rel(lg, ref) < 1e-5


# ---------------------------------------------------------------------------------- fp32-accuracy path kernels
def test_split_and_gemm3_reach_fp32_accuracy(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(21)
    M, N, K = 777, 768, 768
    a = torch.randn(M, K, device=cuda, generator=g)
    w = torch.randn(N, K, device=cuda, generator=g) * K ** -0.5
    bias = torch.randn(N, device=cuda, generator=g)
    resid = torch.randn(M, N, device=cuda, generator=g)
    hi = torch.empty(M, K, device=cuda, dtype=torch.bfloat16); lo = torch.empty_like(hi)
    ops.split_bf16(a, hi, lo, rows=M, D=K)
    assert torch.equal(hi, a.bfloat16()) and _rel(hi.float() + lo.float(), a) < 2e-5
    w_hi = w.bfloat16(); w_lo = (w - w_hi.float()).bfloat16()
    out = torch.empty(M, N, device=cuda)
    ops.gemm3(hi, lo, w_hi, w_lo, out, bias=bias, resid=resid)
    ref = (resid.double() + a.double() @ w.double().t() + bias.double()).float()
    assert _rel(out, ref) < 2e-5


def test_quickgelu_f32_and_embed_assemble(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(22)
    x = torch.randn(1000, 512, device=cuda, generator=g) * 3
    ref = x * torch.sigmoid(1.702 * x)
    ops.quickgelu_f32(x)
    assert _rel(x, ref) < 1e-6
    BT, T, Np, D = 6, 3, 196, 768
    tmp = torch.randn(BT * Np, D, device=cuda, generator=g); pos = torch.randn(Np + 1, D, device=cuda, generator=g)
    te = torch.randn(T, D, device=cuda, generator=g)
    xs = torch.zeros(BT * (Np + 1), D, device=cuda)
    ops.embed_assemble(tmp, xs, pos, te, BT, T, Np)
    want = tmp.view(BT, Np, D) + pos[1:] + te[torch.arange(BT, device=cuda) % T][:, None]
    assert torch.allclose(xs.view(BT, Np + 1, D)[:, 1:], want, atol=1e-6) and (xs.view(BT, Np + 1, D)[:, 0] == 0).all()


@pytest.mark.parametrize("T,N0,H", [(8, 197, 12), (3, 50, 2)])
def test_attention_f32(cuda, T, N0, H):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(23 + T)
    B, G = 2, 8
    D = H * 64
    qkv = torch.randn(B * T * N0, 3 * D, device=cuda, generator=g)
    kv_aux = torch.randn(G + 2 * B * T, 2 * D, device=cuda, generator=g)
    out = torch.zeros(B * T * N0, D, device=cuda)
    ops.attention_f32(qkv, kv_aux, out, B=B, T=T, N0=N0, G=G, H=H)
    assert _rel(out, _attention_ref(qkv, kv_aux, B, T, N0, G, H)) < 1e-5
    # summary form: plain self-attention over the T tokens of every video
    sq = torch.randn(B * T, 3 * D, device=cuda, generator=g)
    so = torch.zeros(B * T, D, device=cuda)
    ops.attention_f32(sq, None, so, B=B, T=1, N0=T, G=0, H=H)
    q, k, v = sq.view(B, T, 3, H, 64).unbind(2)
    aff = torch.einsum("nqhc,nkhc->nqkh", q / 8.0, k).softmax(dim=-2)
    assert _rel(so, torch.einsum("nqlh,nlhc->nqhc", aff, v).reshape(B * T, D)) < 1e-5


def test_prompt_rows_f32_and_patch_gather_split(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(24)
    B, T, G, D = 3, 8, 8, 768
    gp = torch.randn(G, D, device=cuda, generator=g); lp = torch.randn(1, T, D, device=cuda, generator=g)
    cp = torch.randn(B * T, D, device=cuda, generator=g); sm = torch.randn(B * T, D, device=cuda, generator=g)
    w = torch.randn(D, device=cuda, generator=g); b = torch.randn(D, device=cuda, generator=g)
    out = torch.empty(G + 2 * B * T, D, device=cuda)
    ops.prompt_rows_f32(gp, lp, cp, sm, w, b, out, B, T)
    rows = torch.cat([gp, (lp.expand(B, -1, -1) + cp.view(B, T, D)).reshape(B * T, D), sm])
    assert _rel(out, F.layer_norm(rows, (D,), w, b, 1e-5)) < 1e-5
    v = torch.randn(2, 3, 3, 56, 56, device=cuda, generator=g)
    hi = torch.full((2 * 3 * 16, 640), 7.0, device=cuda, dtype=torch.bfloat16); lo = torch.full_like(hi, 7.0)
    ops.patch_gather_split(v, hi, lo, 14)
    ref = F.unfold(v.permute(0, 2, 1, 3, 4).flatten(0, 1), kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert _rel(hi[:, :588].float() + lo[:, :588].float(), ref) < 2e-5 and (hi[:, 588:] == 0).all() and (lo[:, 588:] == 0).all()


@pytest.mark.parametrize("P,H,Kpad", [(16, 224, 768), (14, 56, 640)])
def test_patch_gather_u8(cuda, P, H, Kpad):
    """uint8 HWC frames -> (u/255 - mean)/std -> bf16 patch rows, bit-identical to normalising on the host in fp32."""
    from gava_clip_b200 import ops
    g = torch.Generator().manual_seed(P)
    B, T = 2, 3
    u8 = torch.randint(0, 256, (B, T, H, H, 3), dtype=torch.uint8, generator=g).to(cuda)
    mean, std = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
    gh = H // P
    hi = torch.full((B * T * gh * gh, Kpad), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.patch_gather_u8(u8, hi, None, mean, std, P)
    x = (u8.float() / 255.0 - torch.tensor(mean, device=cuda)) / torch.tensor(std, device=cuda)      # (B,T,H,W,3)
    frames = x.permute(0, 1, 4, 2, 3).flatten(0, 1)
    ref = F.unfold(frames, kernel_size=P, stride=P).transpose(1, 2).reshape(-1, 3 * P * P)
    assert torch.equal(hi[:, :3 * P * P], ref.bfloat16())
    lo = torch.empty_like(hi); hi2 = torch.empty_like(hi)
    ops.patch_gather_u8(u8, hi2, lo, mean, std, P)                                                   # hi/lo (fp32 path)
    assert torch.equal(hi2[:, :3 * P * P], ref.bfloat16())
    assert _rel(hi2[:, :3 * P * P].float() + lo[:, :3 * P * P].float(), ref) < 2e-5


def test_deferred_residual_pair(cuda):
    """LayerNorm(x + delta) and the RESID GEMM epilogue with a second bf16 addend (the deferred attention residual)."""
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(21)
    M, D, K = 1000, 768, 3072
    x = torch.randn(M, D, device=cuda, generator=g)
    delta = torch.randn(M, D, device=cuda, generator=g).bfloat16()
    w = torch.randn(D, device=cuda, generator=g); b = torch.randn(D, device=cuda, generator=g)
    out = torch.empty(M, D, device=cuda, dtype=torch.bfloat16)
    ops.layernorm(x, w, b, out, rows=M, in_row_stride=D, delta=delta)
    ref = F.layer_norm(x + delta.float(), (D,), w, b, 1e-5)
    assert _rel(out.float(), ref) < 1e-2
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    wt = (torch.randn(D, K, device=cuda, generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(D, device=cuda, generator=g)
    want = x + delta.float() + a.float() @ wt.float().t() + bias
    ops.gemm(a, wt, x, mode=ops.EPI_BIAS_RESID_F32, bias=bias, resid=x, aux=delta)
    assert _rel(x, want) < 2e-5


def test_layernorm2_two_stage(cuda):
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(31)
    M, D = 1000, 768
    x = torch.randn(M, D, device=cuda, generator=g) * 3 + 1
    w1, b1, w2, b2 = (torch.randn(D, device=cuda, generator=g) for _ in range(4))
    ref1 = F.layer_norm(x, (D,), w1, b1, 1e-5)
    ref2 = F.layer_norm(ref1, (D,), w2, b2, 1e-5)
    out2 = torch.empty(M, D, device=cuda, dtype=torch.bfloat16)
    ops.layernorm2(x, w1, b1, x, w2, b2, out2, rows=M)          # in place for the f32 output, as the engine uses it
    assert _rel(x, ref1) < 1e-5
    assert _rel(out2.float(), ref2) < 1e-2


@pytest.mark.parametrize("M,D,Kp,N2,step", [(300, 128, 256, 384, 1), (4096 + 77, 768, 3072, 2304, 1), (2560, 1024, 512, 1024, 1),
                                            (197 * 16, 768, 768, 768, 197)])
def test_gemm_layernorm_fold_pair(cuda, M, D, Kp, N2, step):
    """LayerNorm folded around a GEMM pair (GAVA_EPI_BIAS_RESID_STATS_F32 -> GAVA_EPI_LNFOLD_BF16): the producer's x, its
    bf16 copy and row statistics, then  LN(x) W^T + b  from the raw copy -- against an fp32 PyTorch LayerNorm + linear on the
    same x.  step > 1: the consumer reads every step-th row (the cls rows of the last block)."""
    from gava_clip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + D)
    h = torch.randn(M, Kp, device=cuda, generator=g).bfloat16()
    w2 = (torch.randn(D, Kp, device=cuda, generator=g) * Kp ** -0.5).bfloat16()
    b2 = torch.randn(D, device=cuda, generator=g)
    x = torch.randn(M, D, device=cuda, generator=g) * 2.0 + 0.3
    x[:, 5] += 20.0                                    # an outlier channel, as in trained ViTs
    delta = torch.randn(M, D, device=cuda, generator=g).bfloat16()
    x_ref = x + delta.float() + h.float() @ w2.float().t() + b2
    ns = ops.gemm_stats_slots(M, D)
    xb = torch.empty(M, D, device=cuda, dtype=torch.bfloat16)
    stats = torch.full((M, ns, 2), float("nan"), device=cuda)
    ops.gemm(h, w2, x, mode=ops.EPI_BIAS_RESID_STATS_F32, bias=b2, resid=x, aux=delta, copy_out=xb, stats_out=stats)
    assert _rel(x, x_ref) < 2e-5
    assert torch.equal(xb, x.bfloat16())
    assert _rel(stats[:, :, 0].sum(1), x.sum(1)) < 1e-5 and _rel(stats[:, :, 1].sum(1), (x * x).sum(1)) < 1e-5
    # consumer
    gamma = torch.rand(D, device=cuda, generator=g) + 0.5
    beta = torch.randn(D, device=cuda, generator=g) * 0.2
    w = torch.randn(N2, D, device=cuda, generator=g) * D ** -0.5
    b = torch.randn(N2, device=cuda, generator=g)
    wf = (w * gamma[None, :]).bfloat16()
    colsum = wf.float().sum(1).contiguous()
    bf = (b + w @ beta).contiguous()
    rows = torch.arange(0, M, step, device=cuda)
    a = xb.view(M // step, step * D)[:, :D] if step > 1 else xb
    out = torch.empty(rows.numel(), N2, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, wf, out, mode=ops.EPI_LNFOLD_BF16, bias=bf, colsum=colsum, stats_in=stats, stats_row_step=step, eps=1e-5)
    ref = F.layer_norm(x[rows], (D,), gamma, beta, 1e-5) @ w.t() + b
    # the un-folded kernel path for comparison: bf16(LN(x)) @ bf16(w)^T
    plain = F.layer_norm(x[rows], (D,), gamma, beta, 1e-5).bfloat16().float() @ w.bfloat16().float().t() + b
    e_fold, e_plain = _rel(out.float(), ref), _rel(plain.bfloat16().float(), ref)
    # the engine's variant: row-centred weights, no colsum (the mean of x drops out inside the contraction)
    wc = w * gamma[None, :]
    wc = (wc - wc.mean(dim=1, keepdim=True)).bfloat16()
    out_c = torch.empty_like(out)
    ops.gemm(a, wc, out_c, mode=ops.EPI_LNFOLD_BF16, bias=bf, stats_in=stats, stats_row_step=step, eps=1e-5)
    e_cent = _rel(out_c.float(), ref)
    print(f"\n[fold {M}x{D}->{N2}, slots {ns}] folded rel err {e_fold:.2e}, row-centred weights {e_cent:.2e}; "
          f"un-folded bf16 path {e_plain:.2e}")
    assert e_fold < 1e-2 and e_cent < 1e-2
