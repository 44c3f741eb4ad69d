"""One fine-tune step (config-5 model, small batch) for ncu captures: `B=16 python tools/ft_one.py [n_steps]`."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gava_clip_b200 import VitaCLIP, tokenizer
from gava_clip_b200.finetune import FineTuner

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
B, T = int(os.environ.get("B", 16)), 8
clip = torch.randn(B, 3, T, 224, 224, device=dev)
labels = torch.randint(0, 3, (B,), device=dev)
memory = torch.randn(64, 4, 512, device=dev)
ml = torch.randint(0, 3, (64,), device=dev)
tuner = FineTuner(model)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    loss = tuner.step(clip, labels, memory, ml)
torch.cuda.synchronize()
print("loss", float(loss))
