"""N fine-tune steps (config-5 model, small batch) for ncu captures: `B=16 python tools/ft_one.py [n_steps]`."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gava_clip_b200 import tokenizer
from gava_clip_b200.finetune import FineTuner

tokenizer.register_tokens(json.loads(bench.TOKENS.read_text()))
dev = torch.device("cuda:0")
cfg = dict(bench.CONFIGS["cfg2"])
cfg["batch"] = int(os.environ.get("B", 16))
model = bench.build_finetune_model(cfg, dev)
B, T = cfg["batch"], cfg["frames"]
clip = torch.randn(B, 3, T, 224, 224, device=dev)
labels = torch.randint(0, 3, (B,), device=dev)
memory = torch.randn(64, 4, 512, device=dev)
ml = torch.randint(0, 3, (64,), device=dev)
tuner = FineTuner(model)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    loss = tuner.step(clip, labels, memory, ml)
torch.cuda.synchronize()
print("loss", float(loss))
