"""Helpers shared by the parity tests (test infrastructure; may import oracle/)."""
import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import fixtures as fx  # noqa: E402


def load_golden(name):
    return torch.load(fx.GOLDEN / f"{name}.pt", map_location="cpu", weights_only=False)


def build_model(name, device="cpu"):
    """Our drop-in VitaCLIP for a fixture case with the deterministic fixture weights."""
    from gava_clip_b200 import VitaCLIP, tokenizer
    tokenizer.register_tokens(fx.load_tokens())
    kw, B = fx.case_kwargs(name)
    ctx = fx.kapt_cwd() if kw.get("text_prompt_init") else contextlib.nullcontext()
    with ctx:
        model = VitaCLIP(**kw)
    fx.fill_params_(model.state_dict().items(), seed=fx.case_seed(name))
    model.prompt_learner.refresh_token_embeddings(model.textual)   # same order as make_golden.py: weights, then snapshot
    model.eval()
    return model.to(device), kw, B


def rel_err(a, ref):
    a, ref = a.double().cpu(), ref.double().cpu()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def rel_l2(a, ref):
    a, ref = a.double().cpu(), ref.double().cpu()
    return float((a - ref).norm() / ref.norm().clamp_min(1e-30))


def top1_matches(logits, ref, tol):
    """top-1 must agree on every sample whose reference margin exceeds 2*tol*max|ref|."""
    logits, ref = logits.double().cpu(), ref.double().cpu()
    top2 = ref.topk(2, dim=-1).values
    margin = top2[:, 0] - top2[:, 1]
    decisive = margin > 2 * tol * ref.abs().max()
    agree = logits.argmax(-1) == ref.argmax(-1)
    return bool((agree | ~decisive).all()), int((~decisive).sum())
