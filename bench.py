#!/usr/bin/env python
"""Benchmark of the hot path: VitaCLIP video-encoder forward + logit head (BASELINE.json metric: clips/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Configurations (BASELINE.json `configs`, per GPU; clips are independent, ranks shard the batch: weak scaling):
  cfg2 (default, the metric's config)  ViT-B/16,  8 frames 224^2, batch 64, 3-class UPDRS prompts      configs[1]
  cfg3                                 ViT-B/16, 16 frames, K400 400-class prompts, batch 32 (256 over 8 GPUs)  configs[2]
  cfg4                                 ViT-L/14, 16 frames 224^2, batch 16                                  configs[3]
  cfg5 (= --workload finetune)         GaVA-CLIP fine-tune step, KAPT + support memory, batch 64            configs[4]
A step = one forward of one batch of synthetic clips (bf16 tensor-core operands / fp32 accumulate, random-init weights).

One JSON line on rank 0:
  value        clips/s, device-timed (CUDA events), inputs already resident in HBM, max over ranks
  e2e          clips/s through `ClipStreamer(model).run(pinned host batches)`: H2D copy of every clip + D2H read of its
               logits inside the timed region
  roofline     the tcgen05 GEMM kernel (dominant: ~95 % of the FLOPs).  Its share of the step is measured with CUDA
               events around every kernel of 5 consecutive eager steps (after 2 eager warm-up steps); `achieved` =
               algorithmic FLOPs of its launches / (share x the graph-replayed, device-timed ms_per_step), i.e. the
               kernel's time inside the step that `value` reports.  The direct eager-event figure and its step-to-step
               spread are reported beside it.  Fractions against the measured sustained AND burst bf16 peaks.
  finetune     (cfg2 default line only) 5 steps of `FineTuner.step` at batch 64 per GPU -- the path's one collective
               (NCCL SUM all-reduce of the trainable gradients, issued per layer from inside the backward): clips/s,
               ms/step, all-reduce bytes, and the NCCL time left exposed (same steps with the all-reduce switched off)
  cpu_baseline the CPU restatement of the reference algorithm (oracle/, kind "port": the Python reference cannot travel
               to the GPU box) on the box's host cores, bounded sample (batch 1)
`--impl reference` times only that CPU port (all host threads) and prints the same line shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CLASS_DIR = ROOT / "tests" / "golden" / "classes"
TOKENS = ROOT / "tests" / "golden" / "tokens.json"

CONFIGS = {
    "cfg2": dict(arch="b16", frames=8, batch=64, classes="updrs_3cls_classes.txt", n_cls=3,
                 baseline="BASELINE.json configs[1]", metric="clips/sec (8x224^2 ViT-B/16)"),
    "cfg3": dict(arch="b16", frames=16, batch=32, classes="k400_classes.txt", n_cls=400,
                 baseline="BASELINE.json configs[2] (batch 256 over 8 GPUs = 32 per GPU)",
                 metric="clips/sec (16x224^2 ViT-B/16, K400 400-class prompts)"),
    "cfg4": dict(arch="l14", frames=16, batch=16, classes="updrs_3cls_classes.txt", n_cls=3,
                 baseline="BASELINE.json configs[3]", metric="clips/sec (16x224^2 ViT-L/14)"),
}
ARCH_NAME = {"b16": "ViT-B/16", "l14": "ViT-L/14"}


def resolve_config(args) -> dict:
    base = CONFIGS["cfg2" if args.config == "cfg5" else args.config]
    c = dict(base)
    c["name"] = args.config
    for k in ("arch", "frames", "batch"):
        if getattr(args, k) is not None:
            c[k] = getattr(args, k)
    if any(c[k] != base[k] for k in ("arch", "frames", "batch")):
        c["baseline"] = "non-BASELINE shape (flags override the named config)"
        c["metric"] = f"clips/sec ({c['frames']}x224^2 {ARCH_NAME[c['arch']]})"
    c["headline"] = args.config == "cfg2" and all(c[k] == base[k] for k in ("arch", "frames", "batch"))
    return c


def model_kwargs(cfg: dict) -> dict:
    kw = dict(num_frames=cfg["frames"], num_classes=cfg["n_cls"], text_prompt_classes_path=str(CLASS_DIR / cfg["classes"]),
              use_summary_token=True, use_local_prompts=True, use_global_prompts=True, num_global_prompts=8,
              use_text_prompt_learning=True, text_num_prompts=8, text_prompt_CSC=True)
    if cfg["arch"] == "l14":
        kw.update(feature_dim=1024, patch_size=(14, 14), num_heads=16, num_layers=24, embed_dim=768,
                  text_transformer_width=768, text_transformer_heads=12)
    return kw


def tower_config(cfg: dict):
    from gava_clip_b200.config import TowerConfig
    if cfg["arch"] == "l14":
        return TowerConfig(num_frames=cfg["frames"], feature_dim=1024, patch_size=(14, 14), num_heads=16, num_layers=24,
                           embed_dim=768)
    return TowerConfig(num_frames=cfg["frames"])


def load_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class Dist:
    """One process per GPU; NCCL when launched under torchrun."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the sm_100a path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.saved_stdout = None
        if self.world > 1:
            # NCCL prints a "NCCL version ..." banner on the C-level stdout: park fd 1 on stderr until the JSON line is due
            sys.stdout.flush()
            self.saved_stdout = os.dup(1)
            os.dup2(2, 1)
            dist.init_process_group("nccl", device_id=self.dev)
        from gava_clip_b200 import _lib, tokenizer
        _lib.check(_lib.load().gava_init(), "gava_init")
        tokenizer.register_tokens(json.loads(TOKENS.read_text()))      # prompt text -> CLIP ids (no merge table on the GPU box)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def emit(self, line: dict):
        if self.saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(self.saved_stdout, 1)
        print(json.dumps(line), flush=True)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- CPU port (oracle)
def cpu_port_clips_per_s(cfg: dict, steps: int, warmup: int):
    """The oracle (CPU restatement of the reference forward: video tower + text encoder per class + head, as the
    reference's forward does every call) on one clip of the configuration, all host threads."""
    import torch
    from oracle import vita_oracle as vo      # allowed here: this IS the cpu_baseline / reference leg
    from gava_clip_b200 import VitaCLIP, tokenizer
    tokenizer.register_tokens(json.loads(TOKENS.read_text()))
    torch.set_num_threads(os.cpu_count() or 1)
    kw = model_kwargs(cfg)
    model = VitaCLIP(**kw).eval()           # parameter container only; forward is the oracle's
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    T = cfg["frames"]
    clip = torch.randn(1, 3, T, 224, 224, generator=torch.Generator().manual_seed(1234))
    run = lambda: vo.vita_forward(sd, clip, model.tokenized_prompts, num_frames=T, num_heads=kw.get("num_heads", 12),
                                  num_layers=kw.get("num_layers", 12), text_heads=kw.get("text_transformer_heads", 8),
                                  n_ctx=8)
    for _ in range(warmup):
        run()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    return 1.0 / statistics.median(times), statistics.median(times), torch.get_num_threads()


def run_reference(args, cfg):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps, warmup = args.steps, args.warmup
    cps, sec, cores = cpu_port_clips_per_s(cfg, steps, warmup)
    sample = f"{steps} forwards of 1 clip ({cfg['frames']}x224x224, fp32) after {warmup} warm-up, median"
    line = {"impl": "reference", "metric": cfg["metric"], "value": cps, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"VitaCLIP {ARCH_NAME[cfg['arch']]}, {cfg['frames']} frames 224^2, {cfg['n_cls']}-class prompts, "
                                   f"fp32 CPU forward (oracle port of the reference), batch 1 per step"},
            "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- per-kernel event timing
class KernelTimer:
    """CUDA events around every gava_clip_b200.ops call while active (eager mode only: graphs hide the calls)."""

    def __init__(self, names):
        from gava_clip_b200 import ops
        self.ops, self.names, self.records, self.orig = ops, [n for n in names if hasattr(ops, n)], [], {}

    def __enter__(self):
        import torch
        for n in self.names:
            f = getattr(self.ops, n)
            self.orig[n] = f

            def g(*a, _f=f, _n=n, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = _f(*a, **k)
                e1.record()
                fl = 2.0 * a[0].shape[0] * a[0].shape[1] * a[1].shape[0] if _n == "gemm" else 0.0
                tag = f"M{a[0].shape[0]}_N{a[1].shape[0]}_K{a[0].shape[1]}_mode{k.get('mode')}" if _n == "gemm" else None
                self.records.append((_n, e0, e1, fl, tag, torch.cuda.current_stream().cuda_stream))
                return r
            setattr(self.ops, n, g)
        return self

    def __exit__(self, *a):
        for n, f in self.orig.items():
            setattr(self.ops, n, f)

    def drain(self) -> dict:
        """per-kernel {ms, launches, flops} of the records so far (call after a synchronize); clears the records"""
        per, shapes = {}, {}
        main = self.records[0][5] if self.records else None      # the stream the step starts on (patch gather)
        for n, e0, e1, fl, tag, stream in self.records:
            d = per.setdefault(n, {"ms": 0.0, "launches": 0, "flops": 0.0, "main_ms": 0.0, "main_flops": 0.0, "main_launches": 0})
            ms = e0.elapsed_time(e1)
            d["ms"] += ms; d["launches"] += 1; d["flops"] += fl
            if stream == main:         # launches of a side-stream branch overlap the main stream's: not part of its wall time
                d["main_ms"] += ms; d["main_flops"] += fl; d["main_launches"] += 1
            if tag is not None:
                t = shapes.setdefault(tag, {"ms": 0.0, "launches": 0, "flops": 0.0})
                t["ms"] += ms; t["launches"] += 1; t["flops"] += fl
        self.records = []
        self.gemm_shapes = shapes          # per (M, N, K, epilogue mode) of the same records
        return per


FWD_KERNELS = ["gemm", "layernorm", "layernorm2", "patch_gather", "patch_gather_u8", "cls_rows", "temporal_attention",
               "prompt_rows", "attention", "head"]
TRAIN_KERNELS = FWD_KERNELS + ["attention_train", "quickgelu_bf16", "causal_attention_f32", "causal_attention_bwd", "split_bf16",
                               "split3_bf16", "quickgelu_f32", "layernorm_bwd", "attention_bwd", "temporal_attention_bwd",
                               "prompt_rows_bwd", "group_sum", "transpose_bf16", "adamw_step"]


def gemm_roofline(model, clip, peaks, ms_graph_step: float, eager_steps: int = 5):
    """Share of the step per kernel (CUDA events, `eager_steps` consecutive eager steps after 2 eager warm-up steps) and
    the tcgen05 GEMM's achieved TFLOP/s inside the graph-replayed step."""
    import torch
    eng = model.visual._engine
    saved = eng.use_graph
    eng.use_graph = False
    steps = []
    try:
        with torch.no_grad():
            for _ in range(2):
                model(clip)
            walls = []
            with KernelTimer(FWD_KERNELS) as kt:
                for _ in range(eager_steps):
                    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    w0.record()
                    model(clip)
                    w1.record()
                    torch.cuda.synchronize()
                    steps.append(kt.drain())
                    walls.append(w0.elapsed_time(w1))
                gemm_shapes = kt.gemm_shapes
    finally:
        eng.use_graph = saved
    # The temporal-message chain runs on a side-stream branch next to the QKV GEMM, so kernel durations no longer add up
    # to the step: the GEMM's share is (its launches on the MAIN stream, serial among themselves) / (wall time of the eager
    # step), and the FLOPs counted against it are those launches' (the side branch's small GEMMs are 0.03 % of the FLOPs).
    shares, eager_tf, flops = [], [], steps[0]["gemm"]["main_flops"]
    for per, wall in zip(steps, walls):
        shares.append(per["gemm"]["main_ms"] / wall)
        eager_tf.append(per["gemm"]["main_flops"] / (per["gemm"]["main_ms"] * 1e-3) / 1e12)
    share = statistics.median(shares)
    achieved = flops / (share * ms_graph_step * 1e-3) / 1e12
    traffic = None
    tp = ROOT / "profiles" / "gemm_traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
    last = steps[-1]
    tot_last = walls[-1]
    med = statistics.median(eager_tf)
    return {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"],
            "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops_sustained"], "frac_of_burst": achieved / peaks["bf16_tflops"],
            "peak_burst": peaks["bf16_tflops"], "traffic": traffic,
            "traffic_source": "profiles/gemm_traffic.json (ncu --set full: dram__bytes_read+write, mean per launch over the large "
                              "GEMMs -- QKV, out-proj, fc1, fc2 -- of the captured layers; 0.96x their algorithmic operand + output bytes)",
            "peak_source": peaks["_source"] + ": frac vs the sustained bf16 figure (kernel timed inside a long step), frac_of_burst vs the burst figure",
            "how": f"2 MNK of the kernel's main-stream launches / (their median share of the wall time of {eager_steps} event-timed "
                   "eager steps x graph-replayed ms_per_step); kernel_share_of_step = kernel time / eager wall time (side-stream "
                   "launches overlap the main stream, so the shares can sum past 1)",
            "launches_per_step": last["gemm"]["main_launches"], "side_stream_launches_per_step": last["gemm"]["launches"] - last["gemm"]["main_launches"],
            "gflop_per_step": flops / 1e9, "avg_launch_ms": share * ms_graph_step / last["gemm"]["main_launches"],
            "gemm_share_of_step": {"median": share, "min": min(shares), "max": max(shares)},
            "eager_event_tflops": {"median": med, "min": min(eager_tf), "max": max(eager_tf),
                                   "spread": (max(eager_tf) - min(eager_tf)) / med, "stable_within_3pct": (max(eager_tf) - min(eager_tf)) / med < 0.03},
            "kernel_share_of_step": {n: round(d["ms"] / tot_last, 4) for n, d in last.items()},
            # the large GEMM shapes of the last eager step: average launch (us, CUDA events) and TFLOP/s
            "gemm_shapes_eager": {t: {"launches": d["launches"], "avg_us": round(1e3 * d["ms"] / d["launches"], 1),
                                      "tflops": round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 1)}
                                  for t, d in sorted(gemm_shapes.items(), key=lambda kv: -kv[1]["ms"])[:8]}}


# ------------------------------------------------------------------------------------------- fine-tune step (config 5)
def build_finetune_model(cfg, dev):
    import torch
    from gava_clip_b200 import VitaCLIP
    kw = model_kwargs(cfg)
    kw.update(text_prompt_CSC=False, text_prompt_init="cntn_split_uni_disc", knowledge_version=["v0", "v1", "v2", "v3", "v4"],
              cls_type="updrs", use_support_memory=True)
    prev = os.getcwd()
    os.chdir(ROOT / "tests" / "golden" / "kapt_cwd")       # KAPT reads ./data/ke_<type>/ relative to the CWD (kapt_head.py:60)
    try:
        model = VitaCLIP(**kw)
    finally:
        os.chdir(prev)
    with torch.no_grad():                                    # the reference leaves these torch.empty / zero (SURVEY 8a quirks 4,5)
        g0 = torch.Generator().manual_seed(1)
        model.textual.positional_embedding.copy_(torch.randn(model.textual.positional_embedding.shape, generator=g0) * 0.01)
        model.textual.text_projection.copy_(torch.randn(model.textual.text_projection.shape, generator=g0) * 512 ** -0.5)
        for n, p in model.prompt_learner.named_parameters():
            p.copy_(torch.randn(p.shape, generator=g0) * 0.02)
    return model.to(dev).train()


def measure_finetune(D: Dist, cfg: dict, steps: int, warmup: int, detailed: bool) -> dict:
    """BASELINE.json configs[4]: one data-parallel fine-tune step per `step` -- forward with tape, focal-ordinal CE +
    memory loss, backward through the frozen backbone on the sm_100a kernels, NCCL gradient all-reduce (per layer,
    overlapped with the remaining backward), fused AdamW.  KAPT prompt learner (5 knowledge versions) + support memory."""
    import torch
    import torch.distributed as dist
    from gava_clip_b200.finetune import FineTuner
    dev, world, rank = D.dev, D.world, D.rank
    torch.manual_seed(0)
    model = build_finetune_model(cfg, dev)
    B, T = cfg["batch"], cfg["frames"]
    E = model.visual.embed_dim
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    clip = torch.randn(B, 3, T, 224, 224, device=dev, generator=gen)
    labels = torch.randint(0, 3, (B,), device=dev, generator=gen)
    memory = torch.randn(64, 4, E, device=dev, generator=gen)
    mem_labels = torch.randint(0, 3, (64,), device=dev, generator=gen)
    tuner = FineTuner(model, lr=1e-4, weight_decay=0.2)
    n_train = sum(p.numel() for p in tuner.params)

    def timed(n):
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss = tuner.step(clip, labels, memory, mem_labels)
        e1.record()
        D.barrier()
        return D.max_over_ranks(e0.elapsed_time(e1)), loss

    for _ in range(max(warmup, 3)):
        loss = tuner.step(clip, labels, memory, mem_labels)
    eng = model.visual._engine
    launches = eng.launches_last_call + eng.bwd_launches_last_call + 1
    with ClockSampler(D.local) as clocks:
        ms, loss = timed(steps)
    assert torch.isfinite(loss).all()
    out = {"value": B * world * steps / (ms * 1e-3), "unit": "clips/s", "ms_per_step": ms / steps, "steps": steps,
           "warmup": max(warmup, 3), "batch_per_gpu": B, "frames": T, "n_gpus": world, "trainable_params": n_train,
           "allreduce_bytes_per_step": n_train * 4 if world > 1 else 0,
           "collective": f"NCCL SUM all-reduce of {n_train * 4 / 1e6:.0f} MB fp32 gradients per step, one call per layer issued "
                         "from inside the backward + one for the text side / heads" if world > 1 else "none (1 GPU)",
           "gpu_launches_per_step": launches, "loss": float(loss), "clocks": clocks.summary(),
           "workload": "GaVA-CLIP fine-tune step (BASELINE.json configs[4]): KAPT prompt learner (5 knowledge versions) + support "
                       "memory, frozen backbone, focal-ordinal CE, AdamW"}
    if world > 1:
        # every rank applied the same all-reduced gradients: parameters must be bit-identical across ranks
        lo, hi = tuner.flat_p.clone(), tuner.flat_p.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["ranks_bit_identical_after_steps"] = bool(torch.equal(lo, hi))
        # the per-layer overlapped all-reduce must produce the SUM of the ranks' local gradients
        tuner.communicate = False
        tuner.zero_grad(); tuner.forward_backward(clip, labels, memory, mem_labels)
        local_sum = tuner.flat_g.clone()
        dist.all_reduce(local_sum, op=dist.ReduceOp.SUM)
        tuner.communicate = True
        tuner.zero_grad(); tuner.forward_backward(clip, labels, memory, mem_labels); tuner.reduce_gradients()
        torch.cuda.synchronize()
        out["allreduce_vs_sum_of_local_rel_l2"] = float((tuner.flat_g - local_sum).norm() / local_sum.norm())
        # exposed communication: the same steps with the all-reduce switched off
        tuner.communicate = False
        for _ in range(2):
            tuner.step(clip, labels, memory, mem_labels)
        ms_local, _ = timed(steps)
        tuner.communicate = True
        out["ms_per_step_without_allreduce"] = ms_local / steps
        out["nccl_ms_exposed_per_step"] = (ms - ms_local) / steps
    if not detailed:
        del tuner, model
        torch.cuda.empty_cache()
        return out

    # ---- end to end: every step's clip + labels arrive from pinned host memory, the loss is read back every step.
    #   (a) serial: H2D, step, loss D2H, synchronize -- the reference's own loop shape (train.py:421-486)
    #   (b) FineTuner.run(): the H2D copy of batch i+1 overlaps step i
    host_clip = torch.randn(B, 3, T, 224, 224, generator=torch.Generator().manual_seed(99 + rank)).pin_memory()
    dclip = torch.empty_like(clip)
    host_labels = labels.cpu().pin_memory()
    host_loss = torch.empty(()).pin_memory()
    D.barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(steps):
        dclip.copy_(host_clip, non_blocking=True)
        host_loss.copy_(tuner.step(dclip, labels, memory, mem_labels), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e3.record()
    D.barrier()
    ms_serial = D.max_over_ranks(e2.elapsed_time(e3))
    for _ in tuner.run(((host_clip, host_labels) for _ in range(2)), memory, mem_labels):
        pass
    D.barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for hl in tuner.run(((host_clip, host_labels) for _ in range(steps)), memory, mem_labels):
        host_loss = hl
    e5.record()
    D.barrier()
    ms_e2e = D.max_over_ranks(e4.elapsed_time(e5))
    clips_total = B * world * steps
    out["e2e"] = {"value": clips_total / (ms_e2e * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": host_clip.numel() * 4,
                  "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / steps,
                  "api": "gava_clip_b200.finetune.FineTuner.run(pinned host batches): H2D of batch i+1 overlaps step i, loss read back every step",
                  "serial_value": clips_total / (ms_serial * 1e-3), "serial_ms_per_step": ms_serial / steps}
    # ---- per-kernel CUDA-event breakdown of 3 steps (rank 0, single GPU: the eager all-reduce calls are not kernels of ours)
    if rank == 0 and world == 1:
        # eager launches for this part: the timed steps above replay the tower's / text side's CUDA graphs
        eng, tt = model.visual._engine, getattr(model._text_tower, "_trainer", None)
        saved = (eng.use_graph, tt.use_graph if tt is not None else None)
        eng.use_graph = False
        if tt is not None:
            tt.use_graph = False
        try:
            tuner.step(clip, labels, memory, mem_labels)
            with KernelTimer(TRAIN_KERNELS) as kt:
                for _ in range(3):
                    tuner.step(clip, labels, memory, mem_labels)
                torch.cuda.synchronize()
                per = kt.drain()
                out["gemm_shapes"] = {t: {"launches_per_step": d["launches"] // 3, "avg_us": round(1e3 * d["ms"] / d["launches"], 1),
                                          "ms_per_step": round(d["ms"] / 3, 3),
                                          "tflops": round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 1)}
                                      for t, d in sorted(kt.gemm_shapes.items(), key=lambda kv: -kv[1]["ms"])[:14]}
        finally:
            eng.use_graph = saved[0]
            if tt is not None:
                tt.use_graph = saved[1]
        for d in per.values():
            d["ms"] /= 3; d["launches"] //= 3; d["flops"] /= 3
            d["main_ms"] /= 3; d["main_launches"] //= 3; d["main_flops"] /= 3
        out["per_kernel"] = per
    return out


def run_finetune(args, cfg):
    D = Dist()
    ft = measure_finetune(D, cfg, args.steps, args.warmup, detailed=True)
    if D.rank != 0:
        D.close()
        return
    peaks = load_peaks()
    f_clip = tower_config(cfg).flops_per_clip(n_text_rows=15)
    line = {
        "metric": "clips/sec, fine-tune step (fwd + bwd through frozen " + ARCH_NAME[cfg["arch"]] + " + grad all-reduce + AdamW)",
        "value": ft["value"], "unit": "clips/s", "n_gpus": D.world, "steps": args.steps, "warmup": ft["warmup"],
        "ms_per_step": ft["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"{ft['workload']}; {ARCH_NAME[cfg['arch']]}, {cfg['frames']} frames 224^2, batch {cfg['batch']} per GPU",
                   "global_batch": cfg["batch"] * D.world, "trainable_params": ft["trainable_params"],
                   "parallelism": f"dp{D.world} ({ft['collective']})",
                   "l2": "activations (>1 GB per layer) exceed the 126 MB L2; no flush needed"},
        "e2e": ft["e2e"], "gpu_launches": ft["gpu_launches_per_step"] * args.steps, "clocks": ft["clocks"], "loss": ft["loss"],
    }
    for k in ("allreduce_bytes_per_step", "ranks_bit_identical_after_steps", "allreduce_vs_sum_of_local_rel_l2",
              "ms_per_step_without_allreduce", "nccl_ms_exposed_per_step"):
        if k in ft:
            line[k] = ft[k]
    per = ft.get("per_kernel")
    if per:
        total = sum(d["ms"] for d in per.values())
        gm = per["gemm"]
        # main-stream launches only: the chains' small GEMMs run on the side stream next to a capped large GEMM, their
        # event-timed durations include waiting for SMs (and they are 0.1 % of the FLOPs)
        ach = gm["main_flops"] / (gm["main_ms"] * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (fwd + dgrad + wgrad launches)", "achieved": ach,
                            "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"],
                            "frac_of_burst": ach / peaks["bf16_tflops"],
                            "traffic": None, "peak_source": peaks["_source"] + ", sustained bf16 figure",
                            "how": "CUDA events around every kernel of 3 consecutive eager steps; achieved = 2MNK / time of the GEMM "
                                   "launches on the main stream (side-stream launches overlap it: their times in kernel_ms_per_step "
                                   "include waiting for SMs)",
                            "launches_per_step": gm["main_launches"], "side_stream_launches_per_step": gm["launches"] - gm["main_launches"],
                            "kernel_ms_per_step": {n: round(d["ms"], 3) for n, d in sorted(per.items(), key=lambda kv: -kv[1]["ms"])},
                            "kernel_share_of_step": {n: round(d["ms"] / total, 4) for n, d in per.items()},
                            "gemm_shapes_eager": ft.get("gemm_shapes")}
        # forward 1x + backward (dgrad only) ~1x of the forward's dense FLOPs, attention backward 2.5x its forward
        line["step_tensor_util"] = {"algorithmic_gflop_per_clip_fwd": f_clip / 1e9,
                                    "tflops_per_gpu_fwd_equiv_x2": 2 * ft["value"] / D.world * f_clip / 1e12}
    D.emit(line)
    D.close()


# ------------------------------------------------------------------------------------------- our arm (inference)
def run_ours(args, cfg):
    import torch
    from gava_clip_b200 import ClipStreamer, VitaCLIP, tokenizer

    D = Dist()
    dev, world, rank = D.dev, D.world, D.rank
    tokenizer.register_tokens(json.loads(TOKENS.read_text()))
    torch.manual_seed(0)
    model = VitaCLIP(**model_kwargs(cfg)).eval().to(dev)
    B, T = cfg["batch"], cfg["frames"]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    clip = torch.randn(B, 3, T, 224, 224, device=dev, generator=gen)
    host_clip = torch.randn(B, 3, T, 224, 224, generator=torch.Generator().manual_seed(99 + rank)).pin_memory()
    steps, warm = args.steps, max(args.warmup, 3)

    with torch.no_grad():
        # ---------------- device-resident throughput
        model(clip)                                                   # builds workspace + CUDA graph for this shape
        dbuf = model.visual._engine.input_buffer(tuple(clip.shape))  # the graph's static clip buffer (resident in HBM)
        dbuf.copy_(clip)
        for _ in range(warm):
            logits = model(dbuf)[0]
        launches_per_step = model.visual._engine.launches_last_call + 1      # + the head kernel
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(D.local) as clocks:
            e0.record()
            for _ in range(steps):
                logits = model(dbuf)[0]
            e1.record()
            D.barrier()
        ms = D.max_over_ranks(e0.elapsed_time(e1))
        # ---------------- end to end through the public API, clip in PINNED HOST memory, every step:
        #   H2D copy of the step's clip + forward + D2H read of the step's logits, all inside the timed region.
        #   (a) serial: `logits = model(clip.cuda())[0].cpu()` per batch (train.py:654-661)
        #   (b) pipelined: gava_clip_b200.ClipStreamer overlaps the H2D of batch i+1 with the forward of batch i
        host_logits = torch.empty(B, logits.shape[1]).pin_memory()
        for _ in range(2):
            dbuf.copy_(host_clip, non_blocking=True)
            host_logits.copy_(model(dbuf)[0], non_blocking=True)
        D.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(steps):
            dbuf.copy_(host_clip, non_blocking=True)
            host_logits.copy_(model(dbuf)[0], non_blocking=True)
            torch.cuda.current_stream().synchronize()                # the caller reads the logits before the next batch
        e3.record()
        D.barrier()
        ms_serial = D.max_over_ranks(e2.elapsed_time(e3))
        streamer = ClipStreamer(model)
        for out in streamer.run(host_clip for _ in range(3)):        # warm-up (staging buffers, pinned outputs)
            pass
        streamer.h2d_bytes = streamer.d2h_bytes = 0
        D.barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        for out in streamer.run(host_clip for _ in range(steps)):
            host_logits = out
        e5.record()
        D.barrier()
        ms_e2e = D.max_over_ranks(e4.elapsed_time(e5))
        h2d, d2h = streamer.h2d_bytes // steps, streamer.d2h_bytes // steps
        # (c) same, but the clips arrive as the decoder's uint8 frames (B,T,H,W,3): 1/4 of the H2D bytes, normalisation
        #     fused into the patch gather.  Reported as an extra; `e2e.value` stays the fp32-clip call.
        model.visual.set_input_norm([0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711])
        host_u8 = torch.randint(0, 256, (B, T, 224, 224, 3), dtype=torch.uint8,
                                generator=torch.Generator().manual_seed(7 + rank)).pin_memory()
        s8 = ClipStreamer(model)
        for out in s8.run(host_u8 for _ in range(3)):
            pass
        D.barrier()
        e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e6.record()
        for out in s8.run(host_u8 for _ in range(steps)):
            pass
        e7.record()
        D.barrier()
        ms_u8 = D.max_over_ranks(e6.elapsed_time(e7))
        assert torch.isfinite(host_logits).all()

    peaks = load_peaks()
    f_clip = tower_config(cfg).flops_per_clip(n_text_rows=cfg["n_cls"])
    clips_total = B * world * steps
    value = clips_total / (ms * 1e-3)
    roof = gemm_roofline(model, clip, peaks, ms / steps) if rank == 0 else None
    used_graph = bool(model.visual._engine.use_graph)
    del streamer, s8
    model.visual._engine = None             # drop workspaces / graphs before the fine-tune model builds its tape
    del model
    torch.cuda.empty_cache()
    ft = None
    if cfg["headline"] and not args.no_finetune:
        ft = measure_finetune(D, cfg, 5, 3, detailed=False)
    if rank != 0:
        D.close()
        return
    step_tf = value / world * f_clip / 1e12
    line = {
        "metric": cfg["metric"], "value": value, "unit": "clips/s", "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"VitaCLIP {ARCH_NAME[cfg['arch']]}, {T} frames 224^2, batch {B} per GPU, bf16 inference, "
                               f"{cfg['n_cls']}-class text prompts ({cfg['baseline']})", "name": cfg["name"],
                   "global_batch": B * world, "frames": T,
                   "parallelism": f"dp{world} (clips sharded, no collective)",
                   "l2": f"inputs ({clip.numel() * 4 / 1e6:.0f} MB clip) and activations (>1 GB per layer) exceed the 126 MB L2; no flush needed",
                   "cuda_graph": used_graph},
        "e2e": {"value": clips_total / (ms_e2e * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / steps,
                "api": "gava_clip_b200.ClipStreamer(model).run(pinned host batches): H2D of batch i+1 overlaps forward of batch i",
                "serial_value": clips_total / (ms_serial * 1e-3), "serial_ms_per_step": ms_serial / steps,
                "uint8_value": clips_total / (ms_u8 * 1e-3), "uint8_h2d_bytes_per_step": host_u8.numel()},
        "gpu_launches": launches_per_step * steps,
        "clocks": clocks.summary(),
        "roofline": roof,
        "step_tensor_util": {"algorithmic_gflop_per_clip": f_clip / 1e9, "tflops_per_gpu": step_tf,
                             "frac_of_measured_sustained": step_tf / peaks["bf16_tflops_sustained"],
                             "frac_of_measured_burst": step_tf / peaks["bf16_tflops"], "frac_of_nominal_2250": step_tf / 2250.0},
    }
    if ft is not None:
        line["finetune"] = ft
    if world == 1 and not args.no_cpu_baseline:
        n = 5 if cfg["n_cls"] <= 16 else 1          # the reference re-runs the text encoder per class in every forward
        cps, sec, cores = cpu_port_clips_per_s(cfg, n, 1)
        line["cpu_baseline"] = {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port",
                                "sample": f"{n} forward(s) of 1 clip ({T}x224x224, fp32) after 1 warm-up, median {sec:.3f} s"}
    D.emit(line)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configs[1..4]; cfg2 is the metric's config, cfg5 = --workload finetune")
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU per step (overrides the config)")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--arch", default=None, choices=["b16", "l14"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-finetune", action="store_true", help="skip the fine-tune sub-measurement of the default line")
    ap.add_argument("--workload", default="inference", choices=["inference", "finetune"],
                    help="inference = the headline metric; finetune = configs[4] as its own line")
    args = ap.parse_args()
    if args.workload == "finetune":
        args.config = "cfg5"
    cfg = resolve_config(args)
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.config == "cfg5":
        run_finetune(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
