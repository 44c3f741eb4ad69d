"""Generate tests/golden/* by RUNNING THE REFERENCE ITSELF (CPU, fp32) in the build container.

    python oracle/make_golden.py [case ...]

The reference (/root/reference, read-only, never copied) is imported with two sys.modules shims for the
missing `ftfy` / `av` packages.  Per case this writes
    tests/golden/<case>.pt     outputs: logits, logits_mt, logits_vm, video_features, summary, text_features,
                               per-layer cls rows; plus the reference's state_dict key -> shape listing
    tests/golden/tokens.json   prompt text -> CLIP token ids (so the GPU box needs no merge table)
Weights/inputs are regenerated from seeds by oracle/fixtures.py and are not stored.
This script cannot run on the GPU box (no /root/reference there); its outputs travel instead.
"""
from __future__ import annotations

import json
import sys
import time
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import fixtures as fx  # noqa: E402

REF = Path("/root/reference")


def import_reference():
    sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda s: s))
    sys.modules.setdefault("av", types.ModuleType("av"))
    for p in (str(REF), str(REF / "training")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import VitaCLIP_model as ref_model  # noqa
    return ref_model


def write_support_files():
    """class lists are copied as test inputs (plain label lists); KAPT knowledge files are synthetic."""
    fx.CLASSES.mkdir(parents=True, exist_ok=True)
    for n in ("updrs_3cls_classes.txt", "k400_classes.txt"):
        (fx.CLASSES / n).write_text((REF / "classes" / n).read_text())
    ke = fx.KE_ROOT / "data" / "ke_updrs"
    ke.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(0)
    words = ["slow", "shuffling", "steps", "reduced", "arm", "swing", "stooped", "posture", "normal", "stride",
             "freezing", "of", "gait", "short", "festinating", "walk", "with", "mild", "moderate", "severe"]
    for k in range(5):
        np.save(ke / f"EntityEmb_v{k}.npy", rng.standard_normal((3, 768)).astype(np.float32))
        lines = [" ".join(rng.choice(words, size=6 + k + c)) for c in range(3)]
        (ke / f"simQdesc_v{k}.txt").write_text("\n".join(lines) + "\n")


def run_case(ref_model, name: str, tokens: dict) -> None:
    kw, B = fx.case_kwargs(name)
    t0 = time.time()
    kapt = bool(kw.get("text_prompt_init"))
    ctx = fx.kapt_cwd() if kapt else None
    if ctx:
        ctx.__enter__()
    try:
        model = ref_model.VitaCLIP(**kw)
    finally:
        if ctx:
            ctx.__exit__()
    fx.fill_params_(model.state_dict().items(), seed=fx.case_seed(name))      # state_dict tensors alias the parameters
    # the reference snapshots the prefix/suffix token embeddings at construction (text_encoder.py:283-300), i.e.
    # from the backbone loaded BEFORE the prompt learner is built; emulate that order for the fixture weights.
    pl = model.prompt_learner
    with torch.no_grad():
        for c, tok in enumerate(pl.tokenized_prompts):
            emb = model.textual.token_embedding(tok)
            pl.token_prefix[c] = emb[:, :1, :]
            pl.token_suffix[c] = emb[:, 1:-pl.n_ctx, :] if pl.knowledge_aware_prompt else emb[:, 1 + pl.n_ctx:, :]
    model.eval()
    inp = fx.make_inputs(name)
    taps = {"layer_cls": [], "layer_summary": []}
    hooks = []

    def tap_block(m, i, o):          # hooks must return None (a non-None return replaces the output)
        taps["layer_cls"].append(o[0][:, 0].clone())
        if o[1] is not None:
            taps["layer_summary"].append(o[1].clone())

    feats = {}

    def tap_visual(m, i, o):
        feats.update(video_features=o[0].clone(), summary=o[1].clone() if o[1] is not None else None)

    for blk in model.visual.blocks:
        hooks.append(blk.register_forward_hook(tap_block))
    hooks.append(model.visual.register_forward_hook(tap_visual))
    with torch.no_grad():
        logits, logits_mt, logits_vm = model(inp["video"], memory=inp.get("memory"), video_nte=inp.get("video_nte"))
        text_features = model.text_features.clone()
        # desc_wise leaves `model.text_features` = the LAST class's per-description features (reference :272-276,293)
        desc = model(inp["video"], desc_wise=True)[0] if name in ("kapt_b16_t8", "tiny") else None
        desc_text_features = model.text_features.clone() if desc is not None else None
    for h in hooks:
        h.remove()
    n_layers = len(model.visual.blocks)
    out = {
        "case": name, "logits": logits, "logits_mt": logits_mt, "logits_vm": logits_vm,
        "video_features": feats["video_features"], "summary": feats["summary"],
        "text_features": text_features, "desc_text_features": desc_text_features,
        # per-layer taps only for the two small cases (keeps the fixtures small)
        "layer_cls": torch.stack(taps["layer_cls"][:n_layers]) if name in ("tiny", "b16_t8") else None,
        "layer_summary": torch.stack(taps["layer_summary"][:n_layers]) if name in ("tiny", "b16_t8") else None,
        "prompt_flags": {k: kw.get(k, False) for k in ("use_summary_token", "use_local_prompts", "use_global_prompts")},
        "desc_wise": desc,
        "state_dict_shapes": {k: list(v.shape) for k, v in model.state_dict().items()},
        "requires_grad": sorted(n for n, p in model.named_parameters() if p.requires_grad),
        "tokenized_prompts": [t.clone() for t in model.tokenized_prompts],
        "torch": torch.__version__,
    }
    torch.save(out, fx.GOLDEN / f"{name}.pt")
    # prompt text -> ids, recovered from the reference's own tokenisation
    from VitaCLIP_text_encoder_utils import SimpleTokenizer
    tk = SimpleTokenizer()
    classes = [c.replace("_", " ") for c in open(kw["text_prompt_classes_path"]).read().strip().split("\n") if c[0] != "*"]
    if kapt:
        texts = [d + " " + classes[c] for c in range(len(classes)) for d in model.prompt_learner.context_prompt_learner.cls_disc[c]]
    else:
        texts = [" ".join(["X"] * 8) + " " + c + "." for c in classes]
    for t in texts + classes:
        tokens[t] = [49406] + tk.encode(t) + [49407]
    print(f"{name}: logits {tuple(logits.shape)} max|l| {logits.abs().max():.4f}  [{time.time() - t0:.1f}s]")


def main():
    names = sys.argv[1:] or list(fx.CASES)
    write_support_files()
    ref_model = import_reference()
    tok_path = fx.GOLDEN / "tokens.json"
    tokens = json.loads(tok_path.read_text()) if tok_path.exists() else {}
    for n in names:
        run_case(ref_model, n, tokens)
    tok_path.write_text(json.dumps(tokens))


if __name__ == "__main__":
    main()
