"""Shared test fixtures: parity cases, deterministic weights and inputs.  TEST INFRASTRUCTURE ONLY
(imported by tests/, __graft_entry__.smoke(), bench.py's reference/cpu_baseline legs and oracle/make_golden.py).

Weights are not shipped: `fill_params_` writes the same pseudo-random values into ANY module/state_dict with
the reference's parameter names and shapes (per-tensor CPU generator seeded by crc32(name)), so the reference
model here and the sm_100a model on the GPU box hold bit-identical fp32 weights as long as the state_dict
contract holds.  Only outputs (a few KB) are committed under tests/golden/.
"""
from __future__ import annotations

import json
import os
import zlib
from pathlib import Path
from typing import Dict, Iterable, Tuple

import torch

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
CLASSES = GOLDEN / "classes"
KE_ROOT = GOLDEN / "kapt_cwd"          # contains data/ke_updrs/... ; KAPT cases chdir here

_PROMPT_FLAGS = dict(use_summary_token=True, use_local_prompts=True, use_global_prompts=True, num_global_prompts=8,
                     use_text_prompt_learning=True, text_num_prompts=8, text_prompt_CSC=True)

# name -> (VitaCLIP kwargs, batch size).  Shapes follow BASELINE.json configs 1-5 (+ a tiny case for fast CPU tests).
CASES: Dict[str, Tuple[dict, int]] = {
    "tiny": (dict(input_size=(64, 64), num_frames=3, feature_dim=128, num_heads=2, num_layers=2, embed_dim=128,
                  text_transformer_width=128, text_transformer_heads=2, text_transformer_layers=2,
                  num_classes=3, text_prompt_classes_path="updrs_3cls_classes.txt", **_PROMPT_FLAGS), 2),
    # prompt-flag combinations the reference layer handles independently (utils.py:160-197; global prompts are mandatory)
    "tiny_global_only": (dict(input_size=(64, 64), num_frames=3, feature_dim=128, num_heads=2, num_layers=2, embed_dim=128,
                              text_transformer_width=128, text_transformer_heads=2, text_transformer_layers=2, num_classes=3,
                              text_prompt_classes_path="updrs_3cls_classes.txt",
                              **{**_PROMPT_FLAGS, "use_summary_token": False, "use_local_prompts": False}), 2),
    "tiny_no_local": (dict(input_size=(64, 64), num_frames=3, feature_dim=128, num_heads=2, num_layers=2, embed_dim=128,
                           text_transformer_width=128, text_transformer_heads=2, text_transformer_layers=2, num_classes=3,
                           text_prompt_classes_path="updrs_3cls_classes.txt",
                           **{**_PROMPT_FLAGS, "use_local_prompts": False}), 2),
    "tiny_no_summary": (dict(input_size=(64, 64), num_frames=3, feature_dim=128, num_heads=2, num_layers=2, embed_dim=128,
                             text_transformer_width=128, text_transformer_heads=2, text_transformer_layers=2, num_classes=3,
                             text_prompt_classes_path="updrs_3cls_classes.txt",
                             **{**_PROMPT_FLAGS, "use_summary_token": False}), 2),
    "b16_t8": (dict(num_frames=8, num_classes=3, text_prompt_classes_path="updrs_3cls_classes.txt", **_PROMPT_FLAGS), 2),
    # config-2 shapes at batch 8 (two full 128-row GEMM tiles of cls rows, > 1 wave of attention items per SM)
    "b16_t8_b8": (dict(num_frames=8, num_classes=3, text_prompt_classes_path="updrs_3cls_classes.txt", **_PROMPT_FLAGS), 8),
    "b16_t16_k400": (dict(num_frames=16, num_classes=400, text_prompt_classes_path="k400_classes.txt", **_PROMPT_FLAGS), 1),
    "l14_t16": (dict(num_frames=16, feature_dim=1024, patch_size=(14, 14), num_heads=16, num_layers=24, embed_dim=768,
                     text_transformer_width=768, text_transformer_heads=12, num_classes=3,
                     text_prompt_classes_path="updrs_3cls_classes.txt", **_PROMPT_FLAGS), 1),
    "kapt_b16_t8": (dict(num_frames=8, num_classes=3, text_prompt_classes_path="updrs_3cls_classes.txt",
                         text_prompt_init="cntn_split_uni_disc", knowledge_version=["v0", "v1", "v2", "v3", "v4"],
                         cls_type="updrs", use_support_memory=True, add_nte=True,
                         **{**_PROMPT_FLAGS, "text_prompt_CSC": False}), 2),
}


# weight seed per case.  b16_t8 uses seed 1: with seed 0 the random video/text embeddings happen to be almost orthogonal
# (max|logit| 0.14 instead of the 0.4-0.9 seen for seeds 1-4 and for the reference's own init), which
# makes the "relative to max|logit|" parity metric measure noise.
CASE_SEED = {"b16_t8": 1, "b16_t8_b8": 1}


def case_seed(name: str) -> int:
    return CASE_SEED.get(name, 0)


def case_kwargs(name: str) -> Tuple[dict, int]:
    kw, B = CASES[name]
    kw = dict(kw)
    kw["text_prompt_classes_path"] = str(CLASSES / kw["text_prompt_classes_path"])
    return kw, B


def _gen(seed: int, name: str) -> torch.Generator:
    return torch.Generator().manual_seed((seed * 1_000_003 + zlib.crc32(name.encode())) % (2 ** 63))


@torch.no_grad()
def fill_params_(named: Iterable[Tuple[str, torch.Tensor]], seed: int = 0) -> None:
    """Deterministic, name-keyed fp32 values; magnitudes chosen to keep a random-init transformer well scaled."""
    for name, p in named:
        if not p.dtype.is_floating_point:
            continue
        g = _gen(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        shape = tuple(p.shape)
        r = torch.randn(shape, generator=g, dtype=torch.float32) if p.numel() else torch.empty(shape)
        if name.endswith("logit_scale"):
            v = torch.full(shape, 2.6592600)            # ln(1/0.07)
        elif name.endswith("logit_scale_mt") or name.endswith("logit_scale_vm"):
            v = torch.full(shape, 100.0)
        elif "ln" in name.split(".")[-2:][0] or "norm" in name.split(".")[-2:][0]:
            v = 1.0 + 0.1 * r if leaf == "weight" else 0.05 * r
        elif leaf in ("bias", "in_proj_bias"):
            v = 0.02 * r
        elif p.dim() >= 2 and leaf in ("weight", "in_proj_weight") and "token_embedding" not in name:
            fan_out, fan_in = shape[0], int(torch.tensor(shape[1:]).prod())
            v = r * (2.0 / (fan_in + fan_out)) ** 0.5
        elif leaf in ("proj", "text_projection"):
            v = r * shape[0] ** -0.5
        elif "local_prompts" in name or "global_prompts" in name:
            v = 0.05 * r
        else:                                           # embeddings, cls/pos/time, ctx
            v = 0.02 * r
        p.copy_(v.to(p.dtype))


def make_inputs(name: str, seed: int = 1234) -> Dict[str, torch.Tensor]:
    kw, B = case_kwargs(name)
    g = torch.Generator().manual_seed(seed)
    H, W = kw.get("input_size", (224, 224))
    out = {"video": torch.randn(B, 3, kw["num_frames"], H, W, generator=g)}
    if kw.get("use_support_memory"):
        out["memory"] = torch.randn(B, 4, kw.get("embed_dim", 512), generator=g)
    if kw.get("add_nte"):
        out["video_nte"] = torch.randn(B, 70, kw.get("embed_dim", 512), generator=g)
    return out


def load_tokens() -> Dict[str, list]:
    with open(GOLDEN / "tokens.json") as f:
        return json.load(f)


class kapt_cwd:
    """KAPT reads ./data/ke_<type>/ relative to the CWD (kapt_head.py:60-61)."""

    def __enter__(self):
        self.prev = os.getcwd()
        os.chdir(KE_ROOT)

    def __exit__(self, *a):
        os.chdir(self.prev)
