"""CPU: host-side behaviour of the drop-in module (no kernels run)."""
import json
import os

import pytest
import torch

from tests.util import build_model, fx, rel_err


def test_cpu_input_fails_loudly():
    model, kw, B = build_model("tiny")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(fx.make_inputs("tiny")["video"])


def test_bad_shapes_rejected():
    model, kw, B = build_model("tiny")
    with pytest.raises(ValueError):
        model.visual(torch.zeros(1, 3, 64, 64))


def test_checkpoint_roundtrip_with_ddp_prefix(tmp_path):
    """checkpoint.py:96-105 stores `module.`-prefixed keys; train.py:327-329 strips them and loads strict."""
    model, kw, B = build_model("tiny")
    ckpt = {"model": {"module." + k: v.clone() for k, v in model.state_dict().items()}, "next_step": 3}
    torch.save(ckpt, tmp_path / "c.pth")
    other, _, _ = build_model("tiny")
    with torch.no_grad():
        for p in other.parameters():
            p.zero_()
    sd = {k[len("module."):]: v for k, v in torch.load(tmp_path / "c.pth")["model"].items()}
    other.load_state_dict(sd, strict=True)
    for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a, b), k


def test_tokenizer_matches_fixture():
    from gava_clip_b200 import tokenizer
    try:
        tokenizer.find_vocab()
    except FileNotFoundError:
        pytest.skip("CLIP merge table not present on this box")
    tk = tokenizer.SimpleTokenizer()
    for text, ids in fx.load_tokens().items():
        assert [tokenizer.SOT_ID] + tk.encode(text) + [tokenizer.EOT_ID] == ids, text


def test_text_tower_cache_invalidation():
    model, kw, B = build_model("tiny")
    tt = model._text_tower
    a = tt.features(torch.device("cpu"))
    assert tt.features(torch.device("cpu")) is a                      # cached
    with torch.no_grad():
        model.prompt_learner.ctx.add_(0.1)                            # in-place update bumps _version
    b = tt.features(torch.device("cpu"))
    assert b is not a and not torch.allclose(a[0], b[0])


def test_flops_formula():
    from gava_clip_b200.config import VIT_B16_T8, VIT_B16_T16, VIT_L14_T16
    assert abs(VIT_B16_T8.flops_per_clip() / 1e9 - 298.59) < 0.05
    assert abs(VIT_B16_T16.flops_per_clip() / 1e9 - 606.50) < 0.05
    assert abs(VIT_L14_T16.flops_per_clip() / 1e9 - 2749.72) < 0.05


def test_norm1_fold_packing_is_the_layernorm_identity():
    """Host half of the norm1 fold (TowerEngine.pack): with W'' = W diag(gamma) - rowmean and b' = b + W beta,
    LN(x) W^T + b == rstd * (x W''^T) + b' for ANY x -- the mean of x drops out because the rows of W'' sum to zero.  Checked on
    the packed tensors of the tiny model (bf16 rounding of W'' is the only difference from an fp32 LayerNorm + linear)."""
    import torch.nn.functional as F
    from gava_clip_b200.vision_encoder import TowerEngine, LN_EPS
    model, kw, B = build_model("tiny")
    eng = TowerEngine(model.visual, torch.device("cpu"))
    assert eng.fold_norm1 and eng.defer_resid
    eng.pack()
    P = eng.packed
    assert "b0.qkv.wf" not in P                       # block 0's norm1 comes out of the ln_pre pass
    blk = model.visual.blocks[1]
    wf, bf_ = P["b1.qkv.wf"], P["b1.qkv.bf"]
    D = model.visual.feature_dim
    assert wf.dtype == torch.bfloat16 and wf.shape == (3 * D, D) and bf_.shape == (3 * D,)
    w = torch.cat([blk.attn.q_proj.weight, blk.attn.k_proj.weight, blk.attn.v_proj.weight], 0).detach()
    b = torch.cat([blk.attn.q_proj.bias, blk.attn.k_proj.bias, blk.attn.v_proj.bias], 0).detach()
    # rows of the packed operand sum to (almost) zero: |sum| is bf16 rounding noise, far below the row's 1-norm
    assert float((wf.float().sum(1).abs() / wf.float().abs().sum(1)).max()) < 2e-3
    g = torch.Generator().manual_seed(7)
    x = torch.randn(64, D, generator=g) * 3.0 + 1.5                      # a mean the epilogue never sees
    ref = F.layer_norm(x, (D,), blk.norm1.weight.detach(), blk.norm1.bias.detach(), LN_EPS) @ w.t() + b
    mean, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + LN_EPS)
    folded = rstd * (x @ wf.float().t()) + bf_
    assert rel_err(folded, ref) < 4e-3                                    # bf16 rounding of W'' only
    # the parameters change -> the folded operands are re-packed
    with torch.no_grad():
        blk.norm1.weight.mul_(1.5)
    eng.pack()
    ref2 = F.layer_norm(x, (D,), blk.norm1.weight.detach(), blk.norm1.bias.detach(), LN_EPS) @ w.t() + b
    assert rel_err(rstd * (x @ eng.packed["b1.qkv.wf"].float().t()) + eng.packed["b1.qkv.bf"], ref2) < 4e-3


def test_gemm_stats_slots_host_function():
    """gava_gemm_stats_slots is pure host arithmetic (no GPU): one slot per epilogue-warp column share."""
    from gava_clip_b200 import ops
    assert ops.gemm_stats_slots(100864, 768) == 6          # 256-wide tiles: 128 columns per warp
    assert ops.gemm_stats_slots(65792, 1024) == 8
    assert ops.gemm_stats_slots(512, 768) == 12            # small-M shapes run 128-wide tiles: 64 columns per warp
    assert ops.gemm_stats_slots(102, 128) == 2
    with pytest.raises(ValueError):
        ops.gemm_stats_slots(100, 100)


def test_registered_input_registry():
    """TowerEngine.register_input / unregister_input (host bookkeeping of the graphs captured on caller-owned clip buffers):
    re-registration moves a buffer to the back, the registry is capped (oldest forgotten together with its graph), and
    unregistering drops the graph entry that would otherwise keep the buffer alive."""
    from gava_clip_b200.vision_encoder import TowerEngine
    model, kw, B = build_model("tiny")
    eng = TowerEngine(model.visual, torch.device("cpu"))
    key = lambda b: (b.data_ptr(), tuple(b.shape), b.dtype)
    bufs = [torch.zeros(2, 3, 3, 64, 64) for _ in range(10)]
    for b in bufs[:8]:
        eng.register_input(b)
        eng.graphs[tuple(b.shape) + (b.data_ptr(),)] = ("graph", b, 1)
    eng.register_input(bufs[0])                                   # refresh: now the newest
    eng.register_input(bufs[8]); eng.register_input(bufs[9])      # pushes out the two oldest: bufs[1], bufs[2]
    assert len(eng._registered_inputs) == 8
    assert key(bufs[0]) in eng._registered_inputs and key(bufs[1]) not in eng._registered_inputs
    assert tuple(bufs[1].shape) + (bufs[1].data_ptr(),) not in eng.graphs and tuple(bufs[3].shape) + (bufs[3].data_ptr(),) in eng.graphs
    eng.unregister_input(bufs[3])
    assert key(bufs[3]) not in eng._registered_inputs and tuple(bufs[3].shape) + (bufs[3].data_ptr(),) not in eng.graphs
    eng.unregister_input(bufs[3])                                 # idempotent
