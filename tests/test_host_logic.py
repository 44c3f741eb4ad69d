"""CPU: host-side behaviour of the drop-in module (no kernels run)."""
import json
import os

import pytest
import torch

from tests.util import build_model, fx


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
