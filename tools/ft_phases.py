"""Phase timing of one fine-tune step (CUDA events + host wall clock): where the time outside our kernels goes."""
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

def timed(fn, n=5):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); t0 = time.perf_counter(); ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record(); t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n, t_host / n * 1e3

eng = model.visual._engine
print("full step           gpu %.2f ms  host-issue %.2f ms" % timed(lambda: tuner.step(clip, labels, memory, ml)))
with torch.no_grad():
    print("tower forward_train gpu %.2f ms  host-issue %.2f ms" % timed(lambda: eng.forward_train(clip)))
    dff = torch.randn(B * T, 512, device=dev)
    print("tower backward      gpu %.2f ms  host-issue %.2f ms" % timed(lambda: eng.backward(B, T, dff, None)))
def text_only():
    t_hat, offs, tf, counts = model._text_tower.features(dev, need_grad=True)
    t_hat.sum().backward()
print("text fwd+bwd        gpu %.2f ms  host-issue %.2f ms" % timed(text_only))
def text_ac():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        t_hat, offs, tf, counts = model._text_tower.features(dev, need_grad=True)
    t_hat.float().sum().backward()
print("text fwd+bwd bf16ac gpu %.2f ms  host-issue %.2f ms" % timed(text_ac))
print("zero_grad           gpu %.2f ms  host-issue %.2f ms" % timed(tuner.zero_grad))
print("optimizer           gpu %.2f ms  host-issue %.2f ms" % timed(tuner.optimizer_step))
