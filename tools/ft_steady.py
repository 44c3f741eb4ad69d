"""Steady-state (no syncs inside the loop) cost of the pieces of a fine-tune step, CUDA events over N iterations."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gava_clip_b200 import VitaCLIP, tokenizer
from gava_clip_b200.finetune import FineTuner, finetune_loss
from gava_clip_b200.text_train import TextTrainFunction

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
eng = model.visual._engine
dff = torch.randn(B * T, 512, device=dev)

def loop(name, fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); t_issue = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    print(f"{name:34s} gpu {e0.elapsed_time(e1) / n:7.2f} ms/iter   host issue {t_issue:6.2f} ms/iter")

def tower():
    with torch.no_grad():
        eng.forward_train(clip); eng.backward(B, T, dff, None)
def text():
    t_hat, offs, tf, counts = model._text_tower.features(dev, need_grad=True)
    t_hat.sum().backward()
def fwd_bwd():
    out = model(clip, memory=memory)
    finetune_loss(out[0], labels, out[1], ml).backward()
def repack():
    torch.autograd.graph.increment_version(tuner.params); eng.pack_train()
loop("tower fwd+bwd (direct)", tower)
loop("text fwd+bwd (autograd)", text)
loop("model fwd + loss + backward", fwd_bwd)
loop("re-pack after optimizer", repack)
loop("zero_grad + opt step", lambda: (tuner.zero_grad(), tuner.optimizer_step()))
loop("full FineTuner.step", lambda: tuner.step(clip, labels, memory, ml))
