"""CPU tests of the fine-tune step's host logic: loss restatement vs the reference formula, flat parameter / gradient
buffers, and (gloo, world_size 2) the bucketed gradient all-reduce.  The kernels themselves are covered on the GPU
(tests/test_backward_kernels_gpu.py, tests/test_model_gpu.py)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from gava_clip_b200.finetune import allreduce_flat, finetune_loss, flatten_parameters, focal_ordinal_weight


def _ref_focal_ordinal(y_pred, y_true, gamma=2., alpha=.25, beta=0.2, scale=1.0):
    """independent restatement of training/loss_utils.py:25-44 used as the checker"""
    y_true = F.one_hot(y_true, num_classes=y_pred.shape[-1])
    y_pred = y_pred.softmax(dim=-1)
    ordinal_dist = torch.abs(torch.argmax(y_true, dim=-1) - torch.argmax(y_pred, dim=-1))
    weights = (ordinal_dist / (y_pred.shape[-1] - 1)).float()
    focal = alpha * torch.pow(1 - y_pred, gamma)
    combined = (beta * weights.unsqueeze(-1).expand(-1, y_pred.shape[-1]) + focal) * y_true
    return combined.sum(-1) * scale


def test_loss_matches_reference_formula():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(16, 3, generator=g, requires_grad=True)
    labels = torch.randint(0, 3, (16,), generator=g)
    mt = torch.randn(64, 3, generator=g).log_softmax(-1)
    mtl = torch.randint(0, 3, (64,), generator=g)
    w = focal_ordinal_weight(logits, labels)
    assert torch.allclose(w, _ref_focal_ordinal(logits, labels), atol=1e-7)
    loss = finetune_loss(logits, labels, mt, mtl)
    ref = (F.cross_entropy(logits, labels, reduction="none") * _ref_focal_ordinal(logits, labels)).mean() \
        + (0.1 * F.cross_entropy(mt, mtl, reduction="none")).mean()
    assert torch.allclose(loss, ref, atol=1e-7)
    loss.backward()
    assert logits.grad is not None and torch.isfinite(logits.grad).all()


def test_flatten_parameters_preserves_values_and_aliases():
    g = torch.Generator().manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in [(3, 5), (7,), (), (2, 2, 2)]]
    before = [p.detach().clone() for p in ps]
    flat_p, flat_g = flatten_parameters(ps)
    assert flat_p.numel() % 4 == 0 and flat_p.numel() >= sum(p.numel() for p in ps)
    for p, b in zip(ps, before):
        assert torch.equal(p.detach(), b) and p.requires_grad
    flat_p.mul_(2.0)
    for p, b in zip(ps, before):
        assert torch.equal(p.detach(), 2 * b)              # parameters are views of the flat buffer
    sum((p * p).sum() for p in ps).backward()
    off = 0
    for p in ps:                                            # autograd accumulated INTO the flat gradient buffer
        assert torch.allclose(flat_g[off: off + p.numel()].view(p.shape), 2 * p.detach())
        off += p.numel()


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        allreduce_flat(flat, bucket_elems=256)                                   # 4 buckets, ragged tail
        ret[rank] = bool(torch.equal(flat, torch.arange(1000, dtype=torch.float32) * 3))
    finally:
        dist.destroy_process_group()


def test_two_rank_bucketed_gradient_allreduce():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True


import pytest


@pytest.mark.parametrize("name", ["tiny", "tiny_no_local", "tiny_no_summary", "tiny_global_only"])
def test_tower_gradient_layout_covers_trainable_set(name):
    """TowerTrainer's flat gradient buffer (host logic, no kernels): one view per trainable tower parameter of the
    reference's recipe (VitaCLIP_model.py:230-235), right shapes, no overlap, per-layer slices contiguous (the all-reduce
    units), q|k|v of a summary attention adjacent (one wgrad GEMM fills all three).  The partial prompt-flag configurations
    hold only the tensors they have (utils.py:124-139): no local_prompts / no summary attention / neither."""
    from tests.util import build_model
    from gava_clip_b200.train_engine import TowerTrainer, is_trainable_name
    model, kw, _ = build_model(name)
    enc = model.visual
    eng = TowerTrainer(enc, torch.device("cpu"))
    eng._alloc_grads()
    named = eng.trainable_named()
    assert [n for n, _ in named] == [n for n, p in enc.named_parameters() if p.requires_grad]
    assert all(is_trainable_name(n) for n, _ in named)
    base, end = eng.grad_flat.data_ptr(), eng.grad_flat.data_ptr() + eng.grad_flat.numel() * 4
    spans = []
    for n, p in named:
        g = eng.grads[n]
        assert g.numel() == p.numel() and g.view_as(p).shape == p.shape, n
        assert g.is_contiguous() and base <= g.data_ptr() and g.data_ptr() + g.numel() * 4 <= end, n
        spans.append((g.data_ptr(), g.data_ptr() + g.numel() * 4))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))                    # disjoint
    assert sum(e - s for s, e in spans) == eng.grad_flat.numel() * 4                 # and they tile the whole buffer
    L, D = len(enc.blocks), enc.feature_dim
    assert len(eng.layer_slices) == L
    per_layer = sum(p.numel() for n, p in named if n.startswith("blocks.0."))
    assert all(s.numel() == per_layer and s.is_contiguous() for s in eng.layer_slices)
    assert eng.head_slice.numel() == enc.time_embed.numel() + enc.global_prompts.numel()
    assert eng.head_slice.numel() + L * per_layer == eng.grad_flat.numel()
    assert ("blocks.0.local_prompts" in eng.grads) == bool(kw.get("use_local_prompts"))
    assert ("blocks.0.sqkv.weight" in eng.grads) == bool(kw.get("use_summary_token"))
    if kw.get("use_summary_token"):
        w = eng.grads["blocks.0.sqkv.weight"]
        for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
            v = eng.grads[f"blocks.0.summary_attn_layer.{nm}.weight"]
            assert v.data_ptr() == w.data_ptr() + j * D * D * 4
    if name == "tiny_global_only":
        assert per_layer == 0 and all(s.numel() == 0 for s in eng.layer_slices)     # nothing to all-reduce per layer
    # a backbone weight that requires grad has no backward kernel: refused loudly
    enc.blocks[0].mlp.fc1.weight.requires_grad_(True)
    try:
        eng.check_trainable()
        raise AssertionError("expected NotImplementedError")
    except NotImplementedError:
        pass
