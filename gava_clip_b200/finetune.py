"""Data-parallel fine-tune step (BASELINE config 5; reference: training/train.py:351-367,441-494).

    tuner = FineTuner(model, lr=1e-4, weight_decay=0.2)            # model: gava_clip_b200.VitaCLIP on a CUDA device
    loss = tuner.step(video, labels, memory=memo, memory_labels=mt_labels)

One process per GPU.  A step is: forward (video tower on the sm_100a kernels with an activation tape, head + text side in
PyTorch autograd) -> loss (cross-entropy x focal-ordinal weight + memory_loss_weight * CE of the memory logits, the
reference's default recipe) -> backward (train_engine.TowerTrainer.backward) -> gradient all-reduce -> fused AdamW kernel.

Every trainable parameter of the model lives in ONE flat fp32 buffer (the parameters' .data are re-pointed to views of
it, so state_dict / checkpoints are unaffected), with flat gradient, exp_avg and exp_avg_sq buffers beside it: the
all-reduce is a handful of large NCCL calls over NVLink and the optimizer is one kernel launch (gava_adamw_step).
The tower's gradients are reduced per layer, as the backward retires them, on NCCL's stream -- overlapping the remaining
dgrad work -- the text-side / head gradients once autograd has finished.

The reference wraps the model in DistributedDataParallel + torch.optim.AdamW; that also works with this package (the
backward kernels sit behind torch.autograd), FineTuner is the faster native route to the same update.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops

Tensor = torch.Tensor


def focal_ordinal_weight(logits: Tensor, labels: Tensor, gamma: float = 2.0, alpha: float = 0.25, beta: float = 0.2,
                         scale: float = 1.0) -> Tensor:
    """Per-sample loss weight of training/loss_utils.py:25-44 (categorical_ordinal_focal_weight): alpha (1-p_y)^gamma plus
    beta * |argmax(p) - y| / (n_cls - 1).  Differentiable through the softmax, like the reference (no detach)."""
    p = logits.softmax(dim=-1)
    n = p.shape[-1]
    onehot = F.one_hot(labels, num_classes=n).to(p.dtype)
    dist_w = (labels - p.argmax(dim=-1)).abs().to(p.dtype) / (n - 1)
    combined = (beta * dist_w.unsqueeze(-1) + alpha * (1.0 - p).pow(gamma)) * onehot
    return combined.sum(-1) * scale


def finetune_loss(logits: Tensor, labels: Tensor, logits_mt: Optional[Tensor] = None,
                  memory_labels: Optional[Tensor] = None, *, focal_ordinal: bool = True, beta: float = 0.2,
                  memory_loss_weight: float = 0.1, logits_vm: Optional[Tensor] = None, vnte_loss_weight: float = 0.0
                  ) -> Tensor:
    """loss_tot of train.py:446-474: mean(CE * focal-ordinal weight) [+ w * mean(CE(memory logits))] [- w' * mean diag(vm)]."""
    loss = F.cross_entropy(logits, labels, reduction="none")
    if focal_ordinal:
        loss = loss * focal_ordinal_weight(logits, labels, beta=beta)
    loss = loss.mean()
    if logits_mt is not None and memory_labels is not None:
        loss = loss + (memory_loss_weight * F.cross_entropy(logits_mt, memory_labels, reduction="none")).mean()
    if logits_vm is not None and vnte_loss_weight:
        loss = loss - vnte_loss_weight * torch.diag(logits_vm).mean()
    return loss


def flatten_parameters(params: List[torch.nn.Parameter]) -> Tuple[Tensor, Tensor]:
    """Move `params` into one flat fp32 buffer (values preserved, .data re-pointed to views) and give every parameter a
    .grad that is a view of a second flat buffer.  Returns (flat_param, flat_grad); numel padded to a multiple of 4."""
    dev = params[0].device
    total = sum(p.numel() for p in params)
    padded = (total + 3) // 4 * 4
    flat_p = torch.zeros(padded, device=dev, dtype=torch.float32)
    flat_g = torch.zeros(padded, device=dev, dtype=torch.float32)
    off = 0
    for p in params:
        if p.dtype != torch.float32:
            raise TypeError("fine-tuning keeps fp32 master parameters (the reference's autocast semantics)")
        n = p.numel()
        flat_p[off: off + n].copy_(p.detach().reshape(-1))
        p.data = flat_p[off: off + n].view(p.shape)
        p.grad = flat_g[off: off + n].view(p.shape)
        off += n
    return flat_p, flat_g


def allreduce_flat(flat: Tensor, group=None, bucket_elems: int = 1 << 25) -> None:
    """SUM all-reduce of a flat buffer in buckets (NCCL on GPUs, gloo in the CPU tests); averaging is folded into the
    optimizer's grad_scale.  No-op for a single process."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    handles = [dist.all_reduce(flat[i: i + bucket_elems], op=dist.ReduceOp.SUM, group=group, async_op=True)
               for i in range(0, flat.numel(), bucket_elems)]
    for h in handles:
        h.wait()


class FineTuner:
    def __init__(self, model, lr: float = 1e-4, weight_decay: float = 0.2, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, focal_ordinal: bool = True, ordinal_beta: float = 0.2,
                 memory_loss_weight: float = 0.1, process_group=None,
                 lr_schedule: Optional[Callable[[int], float]] = None, text_tf32: bool = False,
                 loss_fn: Optional[Callable[..., Tensor]] = None):
        self.model = model
        tower_ids = {id(p) for p in model.visual.parameters()}
        tower = [p for p in model.visual.parameters() if p.requires_grad]
        rest = [p for p in model.parameters() if p.requires_grad and id(p) not in tower_ids]
        self.params = tower + rest                    # tower gradients first: the rest is ONE contiguous all-reduce
        self._n_tower = sum(p.numel() for p in tower)
        if not self.params:
            raise ValueError("model has no trainable parameters")
        if self.params[0].device.type != "cuda":
            raise RuntimeError("FineTuner needs the model on a CUDA device (gava_adamw_step has no CPU fallback)")
        self.flat_p, self.flat_g = flatten_parameters(self.params)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.focal_ordinal, self.ordinal_beta, self.memory_loss_weight = focal_ordinal, ordinal_beta, memory_loss_weight
        self.group = process_group
        self.lr_schedule = lr_schedule
        # Only the prompt assembly / KAPT projector / cosine head / memory heads are PyTorch matmuls in a step (the text
        # transformer runs on text_train.TextTrainer); text_tf32=True lets those few use TF32 tensor-core products.
        self.text_tf32 = text_tf32
        # loss_fn(logits, logits_mt, logits_vm, labels, memory_labels) -> scalar: plug in the reference's other criteria
        # (training/loss_utils.py: sigmoid_focal_loss, InfoNCE_loss, ...); default = finetune_loss, the shipped recipe
        self.loss_fn = loss_fn
        self.step_count = 0
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        # False = skip every gradient all-reduce (ranks then diverge): only for measuring how much NCCL time a step
        # leaves exposed (bench.py) and for comparing the overlapped reduction with a plain one
        self.communicate = True

    # ---- per-layer overlap: the tower backward calls this as soon as a contiguous slice of its gradient buffer is final
    def _tower_slice_ready(self, grad_slice: Tensor):
        """async SUM all-reduce of a finished slice; the tower backward waits on the handles before it returns."""
        return dist.all_reduce(grad_slice, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def zero_grad(self) -> None:
        self.flat_g.zero_()
        for p in self.params:       # autograd may have replaced a view (it never does for defined grads, but be safe)
            if p.grad is None or p.grad.data_ptr() < self.flat_g.data_ptr() or \
                    p.grad.data_ptr() >= self.flat_g.data_ptr() + self.flat_g.numel() * 4:
                raise RuntimeError("a parameter's .grad no longer aliases the flat gradient buffer "
                                   "(call FineTuner.zero_grad(), not model.zero_grad(set_to_none=True))")

    def forward_backward(self, video: Tensor, labels: Tensor, memory: Optional[Tensor] = None,
                         memory_labels: Optional[Tensor] = None, video_nte: Optional[Tensor] = None,
                         vnte_loss_weight: float = 0.0, loss_scale: float = 1.0) -> Tensor:
        """Accumulates d(loss * loss_scale) into the flat gradient buffer (train.py's batch_split loop calls this per
        micro-batch with loss_scale = 1 / batch_split); returns the un-scaled loss."""
        self.model.train()
        eng = self.model.visual.engine(video.device)
        eng.grad_ready_hook = self._tower_slice_ready if (self.world > 1 and self.communicate) else None
        tf32_prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32_prev or self.text_tf32
        try:
            logits, logits_mt, logits_vm = self.model(video, memory=memory, video_nte=video_nte)
            if self.loss_fn is not None:
                loss = self.loss_fn(logits, logits_mt, logits_vm, labels, memory_labels)
            else:
                loss = finetune_loss(logits, labels, logits_mt, memory_labels, focal_ordinal=self.focal_ordinal,
                                     beta=self.ordinal_beta, memory_loss_weight=self.memory_loss_weight,
                                     logits_vm=logits_vm, vnte_loss_weight=vnte_loss_weight)
            (loss * loss_scale).backward()
        finally:
            eng.grad_ready_hook = None
            torch.backends.cuda.matmul.allow_tf32 = tf32_prev
        self.last_logits = logits.detach()
        return loss.detach()

    def reduce_gradients(self) -> None:
        """Tower gradients were already SUM-reduced slice by slice inside backward (before autograd added them into the
        flat buffer); reduce the rest: text side, logit_scale, memory / NTE heads."""
        if self.world == 1 or not self.communicate:
            return
        if self._n_tower < self.flat_g.numel():
            allreduce_flat(self.flat_g[self._n_tower:], self.group)

    def _check_param_aliasing(self) -> None:
        lo = self.flat_p.data_ptr()
        hi = lo + self.flat_p.numel() * 4
        for p in self.params:
            if not (lo <= p.data_ptr() < hi):
                raise RuntimeError("a trainable parameter no longer lives in FineTuner's flat buffer (model.to()/.float()/"
                                   ".half() after FineTuner(...) re-allocates parameters): build the FineTuner last")

    def optimizer_step(self) -> None:
        self._check_param_aliasing()
        self.step_count += 1
        lr = self.lr_schedule(self.step_count) if self.lr_schedule is not None else self.lr
        ops.adamw_step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, lr=lr, beta1=self.betas[0],
                       beta2=self.betas[1], eps=self.eps, weight_decay=self.wd, step=self.step_count,
                       grad_scale=1.0 / self.world)
        # the flat kernel writes the parameters' storage directly: bump their version counters so packed bf16 copies and
        # cached text features are refreshed on the next forward
        torch.autograd.graph.increment_version(self.params)

    # ---- checkpointing: same layout as torch.optim.AdamW.state_dict() over `self.params` (one param group), so the
    # reference's checkpoint.save_checkpoint / resume_from_checkpoint (optimizer.state_dict() / load_state_dict()) and
    # checkpoints written with torch.optim.AdamW interoperate
    def state_dict(self) -> Dict:
        state, off = {}, 0
        for i, p in enumerate(self.params):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.exp_avg[off: off + n].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off: off + n].view(p.shape).clone()}
            off += n
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: Dict) -> None:
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(self.params):
            raise ValueError(f"optimizer state holds {len(ids)} parameters, this FineTuner has {len(self.params)}")
        g0 = groups[0]
        self.lr, self.betas, self.eps, self.wd = g0["lr"], tuple(g0["betas"]), g0["eps"], g0["weight_decay"]
        off, steps = 0, set()
        for pid, p in zip(ids, self.params):
            n = p.numel()
            st = sd["state"].get(pid)
            if st is None:                      # parameter never stepped: zero moments
                self.exp_avg[off: off + n].zero_(); self.exp_avg_sq[off: off + n].zero_()
            else:
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state {pid}: shape {tuple(st['exp_avg'].shape)} != parameter {tuple(p.shape)}")
                self.exp_avg[off: off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off: off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
            off += n
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): the fused AdamW keeps one step count")
        self.step_count = steps.pop() if steps else 0

    def run(self, host_batches, memory: Optional[Tensor] = None, memory_labels: Optional[Tensor] = None):
        """Training loop over (clip, labels) batches that live in (pinned) HOST memory -- what the reference's DataLoader
        yields (train.py:395-421): the H2D copy of batch i+1 runs on a copy stream underneath step i (two device staging
        buffers), and every step's loss is returned in pinned host memory.  Yields one loss (0-d pinned tensor) per batch."""
        dev = self.params[0].device
        copy_stream = torch.cuda.Stream(dev)
        stage = [None, None]
        staged = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        losses = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event(), torch.cuda.Event()]

        def upload(k, batch):
            clip, labels = batch
            if stage[k] is None or stage[k][0].shape != clip.shape or stage[k][0].dtype != clip.dtype:
                stage[k] = (torch.empty(clip.shape, device=dev, dtype=clip.dtype), torch.empty(labels.shape, device=dev, dtype=labels.dtype))
                # the caching allocator may hand back a block whose last use on the compute stream is still in flight
                copy_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[k])
                stage[k][0].copy_(clip, non_blocking=True)
                stage[k][1].copy_(labels, non_blocking=True)
                staged[k].record(copy_stream)

        it = iter(host_batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        upload(0, nxt)
        i, pending = 0, None
        while True:
            k = i & 1
            try:
                nxt = next(it)
                upload(k ^ 1, nxt)
                more = True
            except StopIteration:
                more = False
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(staged[k])
            loss = self.step(stage[k][0], stage[k][1], memory, memory_labels)
            consumed[k].record(cur)
            losses[k].copy_(loss, non_blocking=True)
            done[k].record(cur)
            if pending is not None:
                done[pending].synchronize()
                yield losses[pending]
            pending = k
            if not more:
                break
            i += 1
        done[pending].synchronize()
        yield losses[pending]

    def step(self, video: Tensor, labels: Tensor, memory: Optional[Tensor] = None, memory_labels: Optional[Tensor] = None,
             video_nte: Optional[Tensor] = None, vnte_loss_weight: float = 0.0) -> Tensor:
        self.zero_grad()
        loss = self.forward_backward(video, labels, memory, memory_labels, video_nte, vnte_loss_weight)
        self.reduce_gradients()
        self.optimizer_step()
        return loss
