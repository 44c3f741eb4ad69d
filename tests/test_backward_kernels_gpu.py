"""Per-kernel parity of the fine-tune (backward / optimizer) entry points on the B200: each against PyTorch autograd
of the same op in fp32, fed the same bf16-rounded operands (include/gava_sm100.h, "Fine-tune step" section)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    return float((a.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("rows,D,dy_bf16", [(1000, 768, True), (37, 1024, False), (512, 128, True)])
def test_layernorm_bwd(cuda, rows, D, dy_bf16):
    from gava_clip_b200 import ops
    g = _gen(rows + D)
    x = torch.randn(rows, D, device=cuda, generator=g) * 2 + 0.3
    w = torch.randn(D, device=cuda, generator=g)
    b = torch.randn(D, device=cuda, generator=g)
    dy = torch.randn(rows, D, device=cuda, generator=g)
    if dy_bf16:
        dy = dy.bfloat16()
    add1 = torch.randn(rows, D, device=cuda, generator=g)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    F.layer_norm(xr, (D,), wr, br, 1e-5).backward(dy.float())
    dx = add1.clone()
    dxb = torch.empty(rows, D, device=cuda, dtype=torch.bfloat16)
    dw = torch.zeros(D, device=cuda); db = torch.zeros(D, device=cuda)
    ops.layernorm_bwd(dy, x, w, dx, rows=rows, D=D, add1=dx, dx_bf16=dxb, dw=dw, db=db)
    ref = add1 + xr.grad
    assert _rel(dx, ref) < 2e-5
    assert _rel(dxb.float(), ref) < 1e-2
    assert _rel(dw, wr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4


def test_layernorm_bwd_strided_cls_rows(cuda):
    """rows addressed with a pitch (the cls row of every frame inside the [BT*N0, D] stream)."""
    from gava_clip_b200 import ops
    g = _gen(5)
    BT, N0, D = 24, 5, 256
    x = torch.randn(BT * N0, D, device=cuda, generator=g)
    w = torch.randn(D, device=cuda, generator=g)
    dy = torch.randn(BT, D, device=cuda, generator=g).bfloat16()
    dx = torch.zeros(BT * N0, D, device=cuda)
    ops.layernorm_bwd(dy, x, w, dx, rows=BT, D=D, x_stride=N0 * D, dx_stride=N0 * D)
    xr = x.view(BT, N0, D)[:, 0].clone().requires_grad_()
    F.layer_norm(xr, (D,), w, torch.zeros_like(w), 1e-5).backward(dy.float())
    got = dx.view(BT, N0, D)
    assert _rel(got[:, 0], xr.grad) < 2e-5 and float(got[:, 1:].abs().max()) == 0.0


def _attn_ref(qkv, kv_aux, B, T, N0, G, H):
    """fp32 reference of gava_attention_fwd on (requires-grad) fp32 tensors."""
    D = H * 64
    BT = B * T
    q, k, v = qkv.view(BT, N0, 3, H, 64).unbind(2)
    ka, va = kv_aux.view(-1, 2, H, 64).unbind(1)
    outs = []
    for f in range(BT):
        b = f // T
        rows = list(range(G)) + [G + b * T + t for t in range(T)] + [G + BT + f]
        kk = torch.cat([k[f], ka[rows]], 0)
        vv = torch.cat([v[f], va[rows]], 0)
        s = torch.einsum("qhc,khc->hqk", q[f] / 8.0, kk).softmax(-1)
        outs.append(torch.einsum("hqk,khc->qhc", s, vv).reshape(N0, D))
    return torch.cat(outs, 0)


@pytest.mark.parametrize("B,T,N0,G,H", [(2, 8, 197, 8, 12), (1, 3, 17, 8, 2), (2, 4, 50, 8, 3), (1, 70, 197, 8, 2), (1, 16, 257, 8, 2)])
def test_attention_bwd(cuda, B, T, N0, G, H):
    from gava_clip_b200 import ops
    g = _gen(B * 1000 + T * 10 + N0)
    D, BT = H * 64, B * T
    qkv = (torch.randn(BT * N0, 3 * D, device=cuda, generator=g) * 1.5).bfloat16()
    kv_aux = (torch.randn(G + 2 * BT, 2 * D, device=cuda, generator=g) * 1.5).bfloat16()
    dout = torch.randn(BT * N0, D, device=cuda, generator=g).bfloat16()
    qr, ar = qkv.float().requires_grad_(), kv_aux.float().requires_grad_()
    o_ref = _attn_ref(qr, ar, B, T, N0, G, H)
    o_ref.backward(dout.float())
    out = o_ref.detach().bfloat16().contiguous()
    dqkv = torch.empty_like(qkv)
    dkv_aux = torch.empty(G + 2 * BT, 2 * D, device=cuda)
    ops.attention_bwd(qkv, kv_aux, out, dout, dqkv, dkv_aux, B=B, T=T, N0=N0, G=G, H=H)
    # bf16 P / dS operands and bf16 outputs: ~2^-8 relative per element, averaged down by the reductions
    def check(dqkv, dkv_aux):
        for name, got, ref in (("dq", dqkv[:, :D], qr.grad[:, :D]), ("dk", dqkv[:, D:2 * D], qr.grad[:, D:2 * D]),
                               ("dv", dqkv[:, 2 * D:], qr.grad[:, 2 * D:]), ("dk_aux", dkv_aux[:, :D], ar.grad[:, :D]),
                               ("dv_aux", dkv_aux[:, D:], ar.grad[:, D:])):
            assert _rel(got.float(), ref) < 2e-2, name
    check(dqkv, dkv_aux)
    # training forward: same output plus the saved log-sum-exp rows, which the backward then reuses
    out2 = torch.empty_like(out)
    lse = torch.empty(BT, H, N0, device=cuda)
    ops.attention_train(qkv, kv_aux, out2, lse, B=B, T=T, N0=N0, G=G, H=H)
    assert _rel(out2.float(), o_ref.detach()) < 1e-2
    q, k, v = qkv.float().view(BT, N0, 3, H, 64).unbind(2)
    ka = kv_aux.float().view(-1, 2, H, 64)[:, 0]
    for f in (0, BT - 1):
        rows = list(range(G)) + [G + (f // T) * T + t for t in range(T)] + [G + BT + f]
        sc = torch.einsum("qhc,khc->hqk", q[f] / 8.0, torch.cat([k[f], ka[rows]], 0))
        assert float((lse[f] - torch.logsumexp(sc, -1) * 1.4426950408889634).abs().max()) < 2e-3
    dqkv2 = torch.zeros_like(qkv); dkv_aux2 = torch.empty_like(dkv_aux)
    ops.attention_bwd(qkv, kv_aux, out2, dout, dqkv2, dkv_aux2, B=B, T=T, N0=N0, G=G, H=H, lse=lse)
    check(dqkv2, dkv_aux2)
    # two-stage (slot + fixed-order sum) accumulation of the prompt-row gradients: same values, bit-reproducible
    slots = torch.empty(BT * H * (G + T + 1) * 128, device=cuda)
    runs = []
    for _ in range(2):
        dq3 = torch.zeros_like(qkv); da3 = torch.full_like(dkv_aux, float("nan"))
        ops.attention_bwd(qkv, kv_aux, out2, dout, dq3, da3, B=B, T=T, N0=N0, G=G, H=H, lse=lse, scratch=slots)
        check(dq3, da3)
        runs.append((dq3, da3))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    # last-layer form: gradients arrive at the cls query row only
    dout1 = torch.zeros_like(dout)
    dout1.view(BT, N0, D)[:, 0] = dout.view(BT, N0, D)[:, 0]
    qr.grad = None; ar.grad = None
    _attn_ref(qr, ar, B, T, N0, G, H).backward(dout1.float())
    dq4 = torch.zeros_like(qkv); da4 = torch.empty_like(dkv_aux)
    ops.attention_bwd(qkv, kv_aux, out2, dout1, dq4, da4, B=B, T=T, N0=N0, G=G, H=H, lse=lse, scratch=slots, n_q=1)
    check(dq4, da4)


@pytest.mark.parametrize("B,T,H", [(3, 8, 12), (2, 70, 2), (1, 3, 2), (2, 17, 3), (1, 128, 2)])
def test_temporal_attention_bwd(cuda, B, T, H):
    from gava_clip_b200 import ops
    g = _gen(B + T + H)
    D = H * 64
    qkv = torch.randn(B * T, 3 * D, device=cuda, generator=g).bfloat16()
    dout = torch.randn(B * T, D, device=cuda, generator=g).bfloat16()
    qr = qkv.float().requires_grad_()
    q, k, v = qr.view(B, T, 3, H, 64).unbind(2)
    s = torch.einsum("bqhc,bkhc->bhqk", q / 8.0, k).softmax(-1)
    torch.einsum("bhqk,bkhc->bqhc", s, v).reshape(B * T, D).backward(dout.float())
    dqkv = torch.empty_like(qkv)
    ops.temporal_attention_bwd(qkv, dout, dqkv, B, T, H)
    assert _rel(dqkv.float(), qr.grad) < 1e-2


def test_prompt_rows_bwd(cuda):
    from gava_clip_b200 import ops
    g = _gen(9)
    G, B, T, D = 8, 3, 4, 768
    BT = B * T
    gp = torch.randn(G, D, device=cuda, generator=g).requires_grad_()
    lp = torch.randn(1, T, D, device=cuda, generator=g).requires_grad_()
    cp = torch.randn(BT, D, device=cuda, generator=g).requires_grad_()
    sm = torch.randn(BT, D, device=cuda, generator=g).requires_grad_()
    w = torch.randn(D, device=cuda, generator=g); b = torch.randn(D, device=cuda, generator=g)
    dn = torch.randn(G + 2 * BT, D, device=cuda, generator=g).bfloat16()
    ext = torch.randn(BT, D, device=cuda, generator=g)
    rows = torch.cat([gp, (lp.expand(B, T, D) + cp.view(B, T, D)).reshape(BT, D), sm], 0)
    F.layer_norm(rows, (D,), w, b, 1e-5).backward(dn.float())
    dg = torch.empty(G, D, device=cuda); dl = torch.empty(BT, D, device=cuda); ds = torch.empty(BT, D, device=cuda)
    ops.prompt_rows_bwd(dn, gp.detach(), lp.detach(), cp.detach(), sm.detach(), w, dg, dl, ds, ext, B, T)
    assert _rel(dg, gp.grad) < 2e-5 and _rel(dl, cp.grad) < 2e-5 and _rel(ds, sm.grad + ext) < 2e-5
    dlp = torch.empty(T, D, device=cuda)
    ops.group_sum(dl, dlp, outer=B, J=T, inner=1)
    assert _rel(dlp, lp.grad[0]) < 2e-5


def test_group_sum_time_embed_shape_and_bias(cuda):
    from gava_clip_b200 import ops
    g = _gen(11)
    B, T, N0, D = 5, 8, 197, 768
    x = torch.randn(B * T * N0, D, device=cuda, generator=g)
    out = torch.empty(T, D, device=cuda)
    ops.group_sum(x, out, outer=B, J=T, inner=N0)
    assert _rel(out, x.view(B, T, N0, D).sum((0, 2))) < 1e-4
    xb = torch.randn(700, 2304, device=cuda, generator=g).bfloat16()
    ob = torch.empty(1, 2304, device=cuda)
    ops.group_sum(xb, ob, outer=700, J=1, inner=1)
    assert _rel(ob[0], xb.float().sum(0)) < 1e-4


def test_transpose_and_wgrad_gemm(cuda):
    """dW = dY^T X through transpose_bf16 + the tcgen05 GEMM (K = rows, zero-padded to 64)."""
    from gava_clip_b200 import ops
    g = _gen(12)
    R, N, K = 280, 2304, 768          # B=4, T=70 rows; K-dim of the wgrad GEMM padded 280 -> 320
    dy = torch.randn(R, N, device=cuda, generator=g).bfloat16()
    x = torch.randn(R, K, device=cuda, generator=g)
    Rp = (R + 63) // 64 * 64
    dyT = torch.empty(N, Rp, device=cuda, dtype=torch.bfloat16)
    xT = torch.empty(K, Rp, device=cuda, dtype=torch.bfloat16)
    ops.transpose_bf16(dy, dyT); ops.transpose_bf16(x, xT)
    assert torch.equal(dyT[:, :R], dy.t()) and float(dyT[:, R:].abs().max()) == 0.0
    assert torch.equal(xT[:, :R], x.bfloat16().t())
    dw = torch.empty(N, K, device=cuda)
    ops.gemm(dyT, xT, dw, mode=ops.EPI_BIAS_F32)
    assert _rel(dw, dy.float().t() @ x.bfloat16().float()) < 2e-5


def test_gemm_dgelu_epilogue(cuda):
    from gava_clip_b200 import ops
    g = _gen(13)
    M, N, K = 1000, 3072, 768
    dy = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    wT = (torch.randn(N, K, device=cuda, generator=g) * K ** -0.5).bfloat16()    # fc2.weight^T  [4D, D]
    a1 = (torch.randn(M, N, device=cuda, generator=g) * 1.5).bfloat16()
    out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    ops.gemm(dy, wT, out, mode=ops.EPI_DGELU_BF16, aux=a1)
    ar = a1.float().requires_grad_()
    (ar * torch.sigmoid(1.702 * ar)).backward(dy.float() @ wT.float().t())
    assert _rel(out.float(), ar.grad) < 2e-2
    h = torch.empty_like(a1)
    ops.quickgelu_bf16(a1, h)
    assert _rel(h.float(), a1.float() * torch.sigmoid(1.702 * a1.float())) < 1e-2


def test_adamw_matches_torch(cuda):
    from gava_clip_b200 import ops
    g = _gen(14)
    n = 4 * 1001
    p = torch.randn(n, device=cuda, generator=g)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=0.2)
    m = torch.zeros(n, device=cuda); v = torch.zeros(n, device=cuda)
    for step in range(1, 4):
        gr = torch.randn(n, device=cuda, generator=g)
        ref.grad = gr.clone() * 0.5
        opt.step()
        ops.adamw_step(p, gr, m, v, lr=1e-3, weight_decay=0.2, step=step, grad_scale=0.5)
    assert _rel(p, ref.detach()) < 1e-5


@pytest.mark.parametrize("n_seq,L,H", [(15, 77, 8), (3, 20, 2), (2, 128, 1), (4, 16, 2), (2, 1, 1), (1, 97, 3)])
def test_causal_attention_bwd(cuda, n_seq, L, H):
    """Tensor-core causal attention of the text transformer (seq_attention_mma.cu): the 3-term split-bf16 forward keeps
    fp32 accuracy (1e-4 of an fp32 PyTorch evaluation), the bf16 backward the 2e-2 budget of the tower's attention backward."""
    from gava_clip_b200 import ops
    g = _gen(n_seq + L + H)
    D = H * 64
    qkv = torch.randn(n_seq * L, 3 * D, device=cuda, generator=g)
    dout = torch.randn(n_seq * L, D, device=cuda, generator=g).bfloat16()
    qr = qkv.clone().requires_grad_()
    q, k, v = qr.view(n_seq, L, 3, H, 64).unbind(2)
    mask = torch.full((L, L), float("-inf"), device=cuda).triu(1)
    s = (torch.einsum("bqhc,bkhc->bhqk", q / 8.0, k) + mask).softmax(-1)
    ref_out = torch.einsum("bhqk,bkhc->bqhc", s, v).reshape(n_seq * L, D)
    ref_out.backward(dout.float())
    out = torch.full((n_seq * L, D), float("nan"), device=cuda)
    ops.causal_attention_f32(qkv, out, n_seq=n_seq, L=L, H=H)
    e_fwd = _rel(out, ref_out.detach())
    dqkv = torch.full((n_seq * L, 3 * D), float("nan"), device=cuda, dtype=torch.bfloat16)
    ops.causal_attention_bwd(qkv, dout, dqkv, n_seq=n_seq, L=L, H=H)
    e_bwd = _rel(dqkv.float(), qr.grad)
    print(f"\n[causal attention {n_seq}x{L}x{H}] forward rel err {e_fwd:.2e}, backward rel err {e_bwd:.2e}")
    assert e_fwd < 1e-4
    assert e_bwd < 2e-2
