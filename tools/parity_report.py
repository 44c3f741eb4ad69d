"""Logit / feature errors of the bf16 path against the reference goldens, per case: `python tools/parity_report.py`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import build_model, fx, load_golden, rel_err, rel_l2
dev = torch.device("cuda:0")
for name in ["tiny", "b16_t8", "b16_t16_k400", "l14_t16", "kapt_b16_t8"]:
    g = load_golden(name)
    model, kw, B = build_model(name, device=dev)
    inp = {k: v.to(dev) for k, v in fx.make_inputs(name).items()}
    with torch.no_grad():
        logits = model(inp["video"], memory=inp.get("memory"), video_nte=inp.get("video_nte"))[0]
    print(f"{name:14s} logits rel err {rel_err(logits, g['logits']):.3e}  video_features rel-L2 {rel_l2(model.video_features, g['video_features']):.3e}"
          f"  summary rel-L2 {rel_l2(model.summary, g['summary']):.3e}  top1 {'ok' if torch.equal(logits.argmax(-1).cpu(), g['logits'].argmax(-1)) else 'DIFF'}")
