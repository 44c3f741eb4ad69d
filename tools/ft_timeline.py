"""Serialised phase timing of the fine-tune step: for each phase, host issue time (GPU idle at start) and GPU time to drain."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gava_clip_b200 import VitaCLIP, tokenizer
from gava_clip_b200.finetune import FineTuner, finetune_loss

tokenizer.register_tokens(json.loads(bench.TOKENS.read_text()))
dev = torch.device("cuda:0")
kw = bench.model_kwargs(8)
kw.update(text_prompt_CSC=False, text_prompt_init="cntn_split_uni_disc", knowledge_version=["v0", "v1", "v2", "v3", "v4"],
          cls_type="updrs", use_support_memory=True)
os.chdir(bench.ROOT / "tests" / "golden" / "kapt_cwd")
model = VitaCLIP(**kw)
with torch.no_grad():
    g0 = torch.Generator().manual_seed(1)
    model.textual.positional_embedding.copy_(torch.randn(model.textual.positional_embedding.shape, generator=g0) * 0.01)
    model.textual.text_projection.copy_(torch.randn(model.textual.text_projection.shape, generator=g0) * 512 ** -0.5)
    for n, p in model.prompt_learner.named_parameters():
        p.copy_(torch.randn(p.shape, generator=g0) * 0.02)
model = model.to(dev).train()
B, T = int(os.environ.get("B", 64)), 8
clip = torch.randn(B, 3, T, 224, 224, device=dev)
labels = torch.randint(0, 3, (B,), device=dev)
memory = torch.randn(64, 4, 512, device=dev)
ml = torch.randint(0, 3, (64,), device=dev)
tuner = FineTuner(model)
for _ in range(3):
    tuner.step(clip, labels, memory, ml)
torch.cuda.synchronize()

acc = {}
def phase(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    a = acc.setdefault(name, [0.0, 0.0])
    a[0] += (t1 - t0) * 1e3; a[1] += (t2 - t0) * 1e3
    return r

N = 5
for _ in range(N):
    phase("zero_grad", tuner.zero_grad)
    out = phase("forward (model)", lambda: model(clip, memory=memory))
    loss = phase("loss", lambda: finetune_loss(out[0], labels, out[1], ml))
    phase("backward", lambda: loss.backward())
    phase("reduce", tuner.reduce_gradients)
    phase("optimizer", tuner.optimizer_step)
for k, (h, g) in acc.items():
    print(f"{k:18s} host-issue {h / N:7.2f} ms   issue+drain {g / N:7.2f} ms")
eng = model.visual._engine
with torch.no_grad():
    phase("tower fwd only", lambda: eng.forward_train(clip)); acc.pop("tower fwd only")
    for _ in range(N):
        phase("tower fwd only", lambda: eng.forward_train(clip))
        phase("pack (after opt)", lambda: (torch.autograd.graph.increment_version(tuner.params), eng.pack_train()))
        phase("text fwd only", lambda: model._text_tower.features(dev, need_grad=False))
for k in ("tower fwd only", "pack (after opt)", "text fwd only"):
    h, g = acc[k]
    print(f"{k:18s} host-issue {h / N:7.2f} ms   issue+drain {g / N:7.2f} ms")
