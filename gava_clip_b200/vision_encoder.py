"""Drop-in `CLIPVisionEncoder` whose forward runs on hand-written sm_100a kernels.

Interface mirrors the reference module (constructor kwargs, parameter names/shapes, return tuple):
    training/VitaCLIP_vision_encoder.py:18-132   (CLIPVisionEncoder)
    training/VitaCLIP_vision_encoder_utils.py:83-203 (TransformerEncoderLayer), :205-220 (ImagePatchEmbed2D)

The nn.Modules below only *hold* parameters under the reference's state_dict keys; the arithmetic is
done by `TowerEngine`, which keeps a packed bf16 copy of the weights and a per-shape workspace in HBM
and drives the C-ABI kernels (ops.py).  There is no PyTorch/CPU fallback for the forward.

HBM layout (all row-major, one frame f = b*T + t after the other):
    x      fp32 [BT*N0, D]   residual stream, live rows only: [cls | patches] per frame
    xn     bf16 [BT*N0, D]   LayerNorm output / attention output (aliased)
    qkv    bf16 [BT*N0, 3D]  q | k | v of the live rows
    kvaux  bf16 [G+2BT, 2D]  k | v of the K/V-only rows: G global prompts, BT local prompts, BT summary tokens
    h      bf16 [BT*N0, 4D]  QuickGELU(fc1)
Prompt tokens never occupy rows of the residual stream: their query rows are dead in the reference
(utils.py:194-197, vision_encoder.py:121) and only their keys/values are ever needed.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from functools import reduce
from operator import mul
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

Tensor = torch.Tensor
LN_EPS = 1e-5


class QuickGELU(nn.Module):
    """x * sigmoid(1.702 x) -- kept for interface parity (utils.py:18-20); the kernels fuse it into fc1."""

    def forward(self, x: Tensor) -> Tensor:
        return x * torch.sigmoid(1.702 * x)


class LayerNorm(nn.LayerNorm):
    """Parameter holder; fp32-statistics semantics (utils.py:22-28) are implemented in gava_layernorm_fwd."""

    def forward(self, x: Tensor) -> Tensor:
        return F.layer_norm(x.float(), self.normalized_shape, self.weight, self.bias, self.eps).to(x.dtype)


class Attention(nn.Module):
    """Holder of q/k/v/out projections (utils.py:31-59): xavier-uniform weights, zero biases."""

    def __init__(self, q_in_dim: int, k_in_dim: int, v_in_dim: int, qk_proj_dim: int, v_proj_dim: int,
                 num_heads: int, out_dim: int):
        super().__init__()
        self.q_proj = nn.Linear(q_in_dim, qk_proj_dim)
        self.k_proj = nn.Linear(k_in_dim, qk_proj_dim)
        self.v_proj = nn.Linear(v_in_dim, v_proj_dim)
        self.out_proj = nn.Linear(v_proj_dim, out_dim)
        self.num_heads = num_heads
        assert qk_proj_dim % num_heads == 0 and v_proj_dim % num_heads == 0
        for lin in (self.q_proj, self.k_proj, self.v_proj, self.out_proj):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)


def _prompt_init_bound(patch_size, dim: int) -> float:
    return math.sqrt(6.0 / float(3 * reduce(mul, patch_size, 1) + dim))


class TransformerEncoderLayer(nn.Module):
    """Parameter holder for one block; key layout of utils.py:104-139 (29 tensors with all prompt options)."""

    def __init__(self, in_feature_dim: int = 768, qkv_dim: int = 768, num_heads: int = 12, mlp_factor: float = 4.0,
                 mlp_dropout: float = 0.0, act=QuickGELU, use_summary_token: bool = False,
                 use_local_prompts: bool = False, num_frames: int = 8, patch_size: Tuple[int, int] = (16, 16)):
        super().__init__()
        d = in_feature_dim
        self.attn = Attention(d, d, d, qkv_dim, qkv_dim, num_heads, d)
        mlp_dim = round(mlp_factor * d)
        self.mlp = nn.Sequential(OrderedDict([
            ("fc1", nn.Linear(d, mlp_dim)), ("act", act()), ("dropout", nn.Dropout(mlp_dropout)),
            ("fc2", nn.Linear(mlp_dim, d)),
        ]))
        self.norm1 = LayerNorm(d)
        self.norm2 = LayerNorm(d)
        self.use_summary_token = use_summary_token
        self.use_local_prompts = use_local_prompts
        if use_summary_token or use_local_prompts:
            self.cls_proj = nn.Linear(d, d)
            self.num_frames = num_frames
        if use_summary_token:
            self.summary_ln = LayerNorm(d)
            self.summary_attn_layer = Attention(d, d, d, qkv_dim, qkv_dim, num_heads, d)
        if use_local_prompts:
            self.local_prompts = nn.Parameter(torch.zeros(1, num_frames, d))
            bound = _prompt_init_bound(patch_size, d)
            nn.init.uniform_(self.local_prompts.data, -bound, bound)
        for lin in (self.mlp.fc1, self.mlp.fc2):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.normal_(lin.bias, std=1e-6)


class ImagePatchEmbed2D(nn.Module):
    """Holder of the patch conv (utils.py:205-220); executed as patch gather + tcgen05 GEMM."""

    def __init__(self, img_size: int = 224, patch_size: int = 16, in_chans: int = 3, embed_dim: int = 768):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


def _ceil_to(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _clip_dims(video: Tensor) -> Tuple[int, int, int, int]:
    """(B, T, H, W) of a float clip (B,3,T,H,W) or a uint8 decoder-layout clip (B,T,H,W,3)."""
    if video.dtype == torch.uint8:
        B, T, H, W, _ = video.shape
    else:
        B, _, T, H, W = video.shape
    return B, T, H, W


class TowerEngine:
    """Packed weights + workspaces + kernel schedule for one CLIPVisionEncoder on one GPU."""

    def __init__(self, enc: "CLIPVisionEncoder", device: torch.device):
        self.enc = enc
        self.device = device
        self.packed: Dict[str, Tensor] = {}
        self.signature = None
        self.ws: Dict[Tuple[int, int], Dict[str, Tensor]] = {}
        self.graphs: Dict[tuple, Tuple[torch.cuda.CUDAGraph, Tensor, int]] = {}
        self.use_graph = os.environ.get("GAVA_CUDA_GRAPH", "1") != "0"
        # The temporal-message chain (eight latency-bound launches on 512 rows, ~65 us per layer back to back) runs as a
        # side-stream branch of the graph next to the block's QKV GEMM.  The persistent GEMM CTAs own a whole SM each, so
        # the small kernels only co-run if the GEMM leaves SMs free: its grid is capped at (SMs - 16) CTAs.  Measured on two
        # boxes (profiles/r02_chain_overlap.txt): uncapped 0 %, 144: +0.7 %, 140: +0.7 %, 136: +1.6 %, 132: +2.0 %, 128: +1.7 %,
        # 124: +1.4 %, <= 120: slower than serial.  GAVA_OVERLAP_CHAIN=0 restores the serial schedule.
        self.overlap_chain = os.environ.get("GAVA_OVERLAP_CHAIN", "1") == "1"
        n_sm = torch.cuda.get_device_properties(device).multi_processor_count if device.type == "cuda" else 148
        self.qkv_ctas = int(os.environ.get("GAVA_QKV_CTAS", str((n_sm - 16) // 2 * 2)))
        # Deferred attention residual: out_proj writes its result as a bf16 delta (no fp32 read-modify-write of the stream,
        # which made that K=768 GEMM HBM-bound), norm2 normalises x + delta, and the fc2 epilogue -- hidden under a K=3072
        # mainloop -- adds delta together with the MLP output.  Same bytes overall, but none of them on a critical path.
        self.defer_resid = os.environ.get("GAVA_DEFER_RESID", "1") != "0"
        # norm1 of blocks 1..L-1 folded around the (fc2 -> QKV) GEMM pair: fc2's epilogue emits bf16(x) + row statistics,
        # the QKV GEMM contracts the raw copy with W diag(norm1.weight) and normalises in its epilogue (no norm1 pass)
        self.fold_norm1 = os.environ.get("GAVA_FOLD_LN1", "1") != "0"
        self.side_stream = torch.cuda.Stream(device) if device.type == "cuda" else None
        self.launches_last_call = 0      # sm_100a kernels launched (or replayed) by the last run()
        self.pack_moves = 0              # times pack() re-allocated (instead of refreshing in place) a packed operand
        self._registered_inputs: Dict[tuple, bool] = {}   # (data_ptr, shape, dtype) of caller-owned static clip buffers, oldest first
        # parity instrument (tests): a dict here receives, per layer, clones of the cls rows of the residual stream and of the
        # summary tokens -- the tensors the reference's per-block forward hooks see (VitaCLIP_vision_encoder.py:116-121).
        # Only honoured by eager runs (a replayed graph cannot clone per layer).
        self.taps: Optional[Dict[str, list]] = None

    # ------------------------------------------------------------------ weights
    def _sig(self):
        return tuple((p.data_ptr(), p._version) for p in self.enc.parameters())

    def pack(self) -> None:
        """bf16 K-major copies of every GEMM weight; q|k|v concatenated to one [3D, D] operand.

        Re-packing after a parameter update (fine-tuning changes the prompt / summary parameters every step) copies
        only the tensors whose source changed, IN PLACE, so captured CUDA graphs and TMA maps stay valid."""
        sig = self._sig()
        if sig == self.signature:
            return
        e, dev = self.enc, self.device
        bf = lambda t: t.detach().to(dev, torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        P = self.packed
        vers = getattr(self, "_src_versions", {})
        moved = False
        copies = []          # (dst, src) refreshes of existing packed tensors, issued as ONE multi-tensor copy

        def put(name: str, srcs, make) -> None:
            """packed[name] = make() unless none of the source parameters changed since the last pack."""
            nonlocal moved
            key = tuple((t.data_ptr(), t._version) for t in srcs)
            if name in P and vers.get(name) == key:
                return
            old = P.get(name)
            if old is not None and len(srcs) == 1 and old.shape == srcs[0].shape and old.device == srcs[0].device:
                if old.data_ptr() != srcs[0].data_ptr():      # an fp32 "copy" may simply alias the parameter
                    copies.append((old, srcs[0].detach()))
            elif old is not None and len(srcs) > 1 and old.shape[0] == sum(t.shape[0] for t in srcs):
                r = 0
                for t in srcs:                                # concatenated q|k|v blocks: slice-wise converting copies
                    copies.append((old[r: r + t.shape[0]], t.detach()))
                    r += t.shape[0]
            else:
                P[name] = make()
                moved = True
            vers[name] = key

        D = e.feature_dim

        def patch_w():
            w = e.patch_embed.proj.weight.detach().reshape(D, -1)
            wp = torch.zeros(D, _ceil_to(w.shape[1], 64), device=dev, dtype=torch.bfloat16)
            wp[:, : w.shape[1]] = w.to(dev, torch.bfloat16)
            return wp

        def proj_wT():
            pj = torch.zeros(_ceil_to(e.embed_dim, 128), D, device=dev, dtype=torch.bfloat16)   # proj^T, rows padded to the N tile
            pj[: e.embed_dim] = e.proj.detach().t().to(dev, torch.bfloat16)
            return pj

        put("patch.w", [e.patch_embed.proj.weight], patch_w)
        put("patch.b", [e.patch_embed.proj.bias], lambda: f32(e.patch_embed.proj.bias))
        put("cls_token", [e.cls_token], lambda: f32(e.cls_token))
        put("pos_embed", [e.pos_embed], lambda: f32(e.pos_embed))
        put("time_embed", [e.time_embed], lambda: f32(e.time_embed))
        for n in ("ln_pre", "ln_post"):
            put(n + ".w", [getattr(e, n).weight], lambda n=n: f32(getattr(e, n).weight))
            put(n + ".b", [getattr(e, n).bias], lambda n=n: f32(getattr(e, n).bias))
        put("proj.wT", [e.proj], proj_wT)
        put("global_prompts", [e.global_prompts], lambda: f32(e.global_prompts))
        for i, blk in enumerate(e.blocks):
            k = f"b{i}."
            a = blk.attn
            put(k + "qkv.w", [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight],
                lambda a=a: bf(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)))
            put(k + "qkv.b", [a.q_proj.bias, a.k_proj.bias, a.v_proj.bias],
                lambda a=a: f32(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0)))
            lins = [("out", a.out_proj), ("fc1", blk.mlp.fc1), ("fc2", blk.mlp.fc2)]
            norms = ["norm1", "norm2"]
            if e.use_summary_token or e.use_local_prompts:          # utils.py:124-126
                lins.append(("cls", blk.cls_proj))
            if e.use_summary_token:                                  # utils.py:129-134
                lins.append(("sout", blk.summary_attn_layer.out_proj))
                norms.append("summary_ln")
            for nm, lin in lins:
                put(k + nm + ".w", [lin.weight], lambda lin=lin: bf(lin.weight))
                put(k + nm + ".b", [lin.bias], lambda lin=lin: f32(lin.bias))
            for n in norms:
                put(k + n + ".w", [getattr(blk, n).weight], lambda blk=blk, n=n: f32(getattr(blk, n).weight))
                put(k + n + ".b", [getattr(blk, n).bias], lambda blk=blk, n=n: f32(getattr(blk, n).bias))
            if e.use_summary_token:
                sa = blk.summary_attn_layer
                put(k + "sqkv.w", [sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight],
                    lambda sa=sa: bf(torch.cat([sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight], 0)))
                put(k + "sqkv.b", [sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias],
                    lambda sa=sa: f32(torch.cat([sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias], 0)))
            if e.use_local_prompts:                                  # utils.py:137-139
                put(k + "local", [blk.local_prompts], lambda blk=blk: f32(blk.local_prompts))
            if i >= 1 and self.fold_norm1 and self.defer_resid:
                # LN(x) W^T + b == rstd (x - mean) W'^T + (b + W beta), W' = W diag(gamma).  Centring the rows of W'
                # (W'' = W' - rowmean(W')) makes W'' orthogonal to the all-ones vector, so (x - mean) W''^T == x W''^T:
                # the mean drops out inside the contraction and the epilogue is one FMA per element (rstd, bias').
                ws, bs = [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], [a.q_proj.bias, a.k_proj.bias, a.v_proj.bias]
                n1 = blk.norm1

                def folded_w(ws=ws, n1=n1):
                    w = torch.cat(ws, 0).detach().to(dev, torch.float32) * n1.weight.detach().to(dev, torch.float32)[None, :]
                    return bf(w - w.mean(dim=1, keepdim=True))
                put(k + "qkv.wf", ws + [n1.weight], folded_w)
                put(k + "qkv.bf", bs + ws + [n1.bias],
                    lambda ws=ws, bs=bs, n1=n1: (torch.cat(bs, 0).detach().to(dev, torch.float32)
                                                 + torch.cat(ws, 0).detach().to(dev, torch.float32) @ n1.bias.detach().to(dev, torch.float32)).contiguous())
        for key in [key for key in P if key.startswith("time_embed@")]:
            del P[key]                                  # resized copies are rebuilt on demand
        if copies:           # fine-tuning refreshes ~100 small tensors per step: a handful of launches instead of one each
            torch._foreach_copy_([d for d, _ in copies], [s_ for _, s_ in copies])
        self._src_versions = vers
        self.signature = sig
        if moved:
            self.graphs.clear()  # captured graphs hold pointers into the old packed tensors
            self.pack_moves += 1  # TowerTrainer's graphs (kept with the tapes) compare this stamp

    def pack_precise(self) -> Dict[str, Tensor]:
        """hi/lo bf16 pairs of every GEMM weight for the fp32-accuracy path (w = hi + lo up to 2^-17)."""
        if getattr(self, "_precise_sig", None) == self.signature and hasattr(self, "packed_lo"):
            return self.packed_lo
        e, dev = self.enc, self.device
        D = e.feature_dim

        def pair(w32: Tensor, name: str, out: Dict[str, Tensor]):
            w32 = w32.detach().to(dev, torch.float32).contiguous()
            hi = w32.to(torch.bfloat16)
            out[name] = (hi, (w32 - hi.float()).to(torch.bfloat16))

        Q: Dict[str, Tuple[Tensor, Tensor]] = {}
        w = e.patch_embed.proj.weight.detach().reshape(D, -1).to(dev, torch.float32)
        wp = torch.zeros(D, _ceil_to(w.shape[1], 64), device=dev)
        wp[:, : w.shape[1]] = w
        pair(wp, "patch.w", Q)
        pj = torch.zeros(_ceil_to(e.embed_dim, 128), D, device=dev)
        pj[: e.embed_dim] = e.proj.detach().t().to(dev, torch.float32)
        pair(pj, "proj.wT", Q)
        for i, blk in enumerate(e.blocks):
            k = f"b{i}."
            a = blk.attn
            pair(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0), k + "qkv.w", Q)
            pair(a.out_proj.weight, k + "out.w", Q)
            pair(blk.mlp.fc1.weight, k + "fc1.w", Q)
            pair(blk.mlp.fc2.weight, k + "fc2.w", Q)
            if e.use_summary_token or e.use_local_prompts:
                pair(blk.cls_proj.weight, k + "cls.w", Q)
            if e.use_summary_token:
                sa = blk.summary_attn_layer
                pair(torch.cat([sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight], 0), k + "sqkv.w", Q)
                pair(sa.out_proj.weight, k + "sout.w", Q)
        self.packed_lo, self._precise_sig = Q, self.signature
        return Q

    def time_embed_for(self, T: int) -> Tensor:
        """nearest-neighbour resize when T differs from the trained length (vision_encoder.py:91-97)."""
        te = self.packed["time_embed"]
        if te.shape[0] == T:
            return te
        key = f"time_embed@{T}"
        if key not in self.packed:
            r = F.interpolate(te.t().unsqueeze(0), size=(T,), mode="nearest").squeeze(0).t().contiguous()
            self.packed[key] = r
        return self.packed[key]

    # ------------------------------------------------------------------ workspace
    MAX_SHAPES = 4    # workspaces / graphs kept alive (an eval loop sees the full batch size and one ragged tail)

    def _evict(self) -> None:
        """Drop the oldest clip shapes (2.3 GB of workspace each at config 2) once more than MAX_SHAPES were seen."""
        while len(self.ws) > self.MAX_SHAPES:
            old = next(iter(self.ws))
            del self.ws[old]
            for k in [k for k, v in self.graphs.items() if _clip_dims(v[1])[:2] == old[-2:]]:
                del self.graphs[k]

    def workspace(self, B: int, T: int) -> Dict[str, Tensor]:
        key = (B, T)
        if key in self.ws:
            return self.ws[key]
        self._evict()
        e, dev = self.enc, self.device
        D, N0, G = e.feature_dim, int(e.num_patches), e.num_global_prompts
        BT = B * T
        M = BT * N0
        Dm = e.blocks[0].mlp.fc1.out_features
        bf, f32 = torch.bfloat16, torch.float32
        z = lambda *s, dt: torch.empty(*s, device=dev, dtype=dt)
        W = {
            "patches": z(BT * (N0 - 1), self.packed["patch.w"].shape[1], dt=bf),
            "x": z(M, D, dt=f32),
            "xn": z(M, D, dt=bf),
            "delta": z(M, D, dt=bf),
            "qkv": z(M, 3 * D, dt=bf),
            "h": z(M, Dm, dt=bf),
            "cls_bf": z(BT, D, dt=bf),
            "cls_bf2": z(BT, D, dt=bf),       # side-stream copy (cls_bf is used by the main stream in the last layer)
            "cls_proj": z(BT, D, dt=f32),
            "x_cls": z(BT, D, dt=f32),
            "sum_n": z(BT, D, dt=bf),
            "sqkv": z(BT, 3 * D, dt=bf),
            "smix": z(BT, D, dt=bf),
            "summary": z(BT, D, dt=f32),
            "aux_n": z(G + 2 * BT, D, dt=bf),
            "kvaux": z(G + 2 * BT, 2 * D, dt=bf),
            "frame_feat": z(BT, self.packed["proj.wT"].shape[0], dt=f32),
        }
        if self.fold_norm1 and self.defer_resid:
            W["xb"] = z(M, D, dt=bf)                                         # bf16 copy of the stream after each block
            W["xstats"] = z(M, ops.gemm_stats_slots(M, D), 2, dt=f32)       # its per-row (sum, sum of squares) partials
        # inputs of the prompt-row kernel that a disabled prompt group does not produce (zeros: their K/V rows are masked
        # out by the attention kernels' key_groups, they only have to be finite)
        if not e.use_local_prompts:
            W["zero_local"] = torch.zeros(1, T, D, device=dev, dtype=f32)
        if not (e.use_summary_token or e.use_local_prompts):
            W["cls_proj"].zero_()
        if not e.use_summary_token:
            W["summary"].zero_()
        self.ws[key] = W
        return W

    # ------------------------------------------------------------------ CUDA-graph replay
    def input_buffer(self, shape, dtype=torch.float32) -> Tensor:
        """Static clip buffer of the captured graph for this shape: callers that fill it directly (e.g. an
        H2D copy from pinned memory) and then pass it to forward() skip the staging copy."""
        key = tuple(shape)
        ent = self.graphs.get(key)
        if ent is not None:
            return ent[1]
        return torch.empty(*shape, device=self.device, dtype=dtype)

    def register_input(self, buf: Tensor) -> None:
        """Declare `buf` a long-lived device clip buffer (e.g. a streaming loop's staging slot): forwards on it replay a
        graph captured on the buffer itself instead of copying it into the shape's static input first (616 MB of HBM
        traffic per step at config 2).  At most a handful of buffers per shape (each holds its own captured graph)."""
        ent = (buf.data_ptr(), tuple(buf.shape), buf.dtype)
        self._registered_inputs.pop(ent, None)
        self._registered_inputs[ent] = True
        while len(self._registered_inputs) > 8:          # forget the oldest registrations (and the graphs captured on them)
            oldest = next(iter(self._registered_inputs))
            del self._registered_inputs[oldest]
            self.graphs.pop(oldest[1] + (oldest[0],), None)

    def unregister_input(self, buf: Tensor) -> None:
        """Forget a registered buffer and drop the graph captured on it (the graph entry holds a reference to the buffer,
        which would otherwise keep e.g. a finished streaming loop's 308 MB staging slots alive)."""
        ent = (buf.data_ptr(), tuple(buf.shape), buf.dtype)
        if self._registered_inputs.pop(ent, None) is not None:
            self.graphs.pop(ent[1] + (ent[0],), None)

    def run(self, video: Tensor) -> Tuple[Tensor, Tensor]:
        """encode_frames, replayed from a CUDA graph per clip shape (launch-bound inner loop: ~190 launches)."""
        if self.enc.precision == "fp32":
            n0 = ops.launch_count()
            out = self.encode_frames_precise(video)
            self.launches_last_call = ops.launch_count() - n0
            return out
        if not self.use_graph or torch.cuda.is_current_stream_capturing():
            n0 = ops.launch_count()
            out = self.encode_frames(video)
            self.launches_last_call = ops.launch_count() - n0
            return out
        # a registered (long-lived) input buffer gets a graph of its own that reads it in place: no staging copy
        reg = (video.data_ptr(), tuple(video.shape), video.dtype) in self._registered_inputs
        key = tuple(video.shape) + ((video.data_ptr(),) if reg else ())
        ent = self.graphs.get(key)
        if ent is None:
            static_in = video if reg else video.clone()
            self.encode_frames(static_in)                 # warm-up: kernel attributes, workspace, allocator
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g):
                self.encode_frames(static_in)
            ent = (g, static_in, ops.launch_count() - n0)
            self.graphs[key] = ent
        g, static_in, n_launch = ent
        if video.data_ptr() != static_in.data_ptr():
            static_in.copy_(video)
        g.replay()
        self.launches_last_call = n_launch
        B, T, _, _ = _clip_dims(video)
        W = self.ws[(B, T)]
        return W["frame_feat"], (W["summary"] if self.enc.use_summary_token else None)

    # ------------------------------------------------------------------ schedule
    def encode_frames(self, video: Tensor) -> Tuple[Tensor, Tensor]:
        """video fp32 [B,3,T,H,W] -> (frame_feat fp32 [BT, >=E], summary tokens fp32 [BT, D]).

        Everything up to (not including) the temporal mean of vision_encoder.py:128-130.
        """
        e, P = self.enc, self.packed
        B, T, Hh, Ww = _clip_dims(video)
        D, N0, G, H = e.feature_dim, int(e.num_patches), e.num_global_prompts, e.num_heads
        BT = B * T
        M = BT * N0
        W = self.workspace(B, T)
        te = self.time_embed_for(T)
        x, xn, qkv, h = W["x"], W["xn"], W["qkv"], W["h"]

        # patch embedding (+bias +pos +time), cls rows, ln_pre          vision_encoder.py:105-113
        if video.dtype == torch.uint8:     # input stage fused into the gather
            ops.patch_gather_u8(video, W["patches"], None, e.input_norm[0], e.input_norm[1], e.patch_embed.patch_size)
        else:
            ops.patch_gather(video, W["patches"], e.patch_embed.patch_size)
        ops.gemm(W["patches"], P["patch.w"], x, mode=ops.EPI_PATCH_F32, bias=P["patch.b"], pos=P["pos_embed"],
                 time_embed=te, group=N0 - 1, group_stride=N0, row_offset=1, T=T)
        ops.cls_rows(x, N0, P["cls_token"], P["pos_embed"], te, BT, T)
        # ln_pre and block 0's norm1 in one pass (the second LayerNorm works on the values still in registers)
        ops.layernorm2(x, P["ln_pre.w"], P["ln_pre.b"], x, P["b0.norm1.w"], P["b0.norm1.b"], xn, rows=M)

        L = len(e.blocks)
        kg = (1 if e.use_local_prompts else 0) | (2 if e.use_summary_token else 0)      # prompt groups among the keys
        main = torch.cuda.current_stream(self.device)
        side = self.side_stream if self.overlap_chain else None
        norm1_done = True       # True when xn already holds norm1(x) of the coming block (block 0: done by the ln_pre pass)
        fold = self.fold_norm1 and self.defer_resid
        for i in range(L):
            k = f"b{i}."
            folded_in = fold and i >= 1      # this block's input came with its bf16 copy + row statistics (previous fc2)
            # ---- temporal message: cls_proj -> summary attention -> K/V-only prompt rows   utils.py:160-188
            # Eight small launches that depend only on the cls rows of x; they run on a side stream (a parallel branch of
            # the CUDA graph) next to norm1 + the QKV GEMM of the live rows and join before the attention kernel.
            if side is not None:
                ev_x = torch.cuda.Event()
                ev_x.record(main)
                side.wait_event(ev_x)
            with torch.cuda.stream(side if side is not None else main):
                if e.use_summary_token or e.use_local_prompts:                               # utils.py:160-165
                    ops.layernorm(x, None, None, W["cls_bf2"], rows=BT, in_row_stride=N0 * D)
                    ops.gemm(W["cls_bf2"], P[k + "cls.w"], W["cls_proj"], mode=ops.EPI_BIAS_F32, bias=P[k + "cls.b"])
                if e.use_summary_token:                                                      # utils.py:168-172
                    ops.layernorm(W["cls_proj"], P[k + "summary_ln.w"], P[k + "summary_ln.b"], W["sum_n"], rows=BT,
                                  in_row_stride=D)
                    ops.gemm(W["sum_n"], P[k + "sqkv.w"], W["sqkv"], mode=ops.EPI_BIAS_BF16, bias=P[k + "sqkv.b"])
                    ops.temporal_attention(W["sqkv"], W["smix"], B, T, H)
                    ops.gemm(W["smix"], P[k + "sout.w"], W["summary"], mode=ops.EPI_BIAS_RESID_F32, bias=P[k + "sout.b"],
                             resid=W["cls_proj"])
                ops.prompt_rows(P["global_prompts"][i], P[k + "local"] if e.use_local_prompts else W["zero_local"],
                                W["cls_proj"], W["summary"], P[k + "norm1.w"], P[k + "norm1.b"], W["aux_n"], B, T, LN_EPS)
                ops.gemm(W["aux_n"], P[k + "qkv.w"][D:], W["kvaux"], mode=ops.EPI_BIAS_BF16, bias=P[k + "qkv.b"][D:])
                if side is not None:
                    ev_kv = torch.cuda.Event()
                    ev_kv.record(side)
            # ---- x = x + attn(norm1(x)) on the live rows                                     utils.py:190-191
            if not norm1_done and not folded_in:
                ops.layernorm(x, P[k + "norm1.w"], P[k + "norm1.b"], xn, rows=M, in_row_stride=D)
            norm1_done = False
            if i < L - 1:
                if folded_in:
                    ops.gemm(W["xb"], P[k + "qkv.wf"], qkv, mode=ops.EPI_LNFOLD_BF16, bias=P[k + "qkv.bf"],
                             stats_in=W["xstats"], eps=LN_EPS,
                             max_ctas=self.qkv_ctas if side is not None else 0)
                else:
                    ops.gemm(xn, P[k + "qkv.w"], qkv, mode=ops.EPI_BIAS_BF16, bias=P[k + "qkv.b"],
                             max_ctas=self.qkv_ctas if side is not None else 0)
                if side is not None:
                    main.wait_event(ev_kv)
                ops.attention(qkv, W["kvaux"], xn, B=B, T=T, N0=N0, G=G, H=H, key_groups=kg)
                if self.defer_resid:
                    dl = W["delta"]
                    ops.gemm(xn, P[k + "out.w"], dl, mode=ops.EPI_BIAS_BF16, bias=P[k + "out.b"])
                    # ---- x = (x + delta) + fc2(QuickGELU(fc1(norm2(x + delta))))              utils.py:191,199
                    ops.layernorm(x, P[k + "norm2.w"], P[k + "norm2.b"], xn, rows=M, in_row_stride=D, delta=dl)
                    ops.gemm(xn, P[k + "fc1.w"], h, mode=ops.EPI_BIAS_GELU_BF16, bias=P[k + "fc1.b"])
                    if fold:
                        ops.gemm(h, P[k + "fc2.w"], x, mode=ops.EPI_BIAS_RESID_STATS_F32, bias=P[k + "fc2.b"], resid=x, aux=dl,
                                 copy_out=W["xb"], stats_out=W["xstats"])
                    else:
                        ops.gemm(h, P[k + "fc2.w"], x, mode=ops.EPI_BIAS_RESID_F32, bias=P[k + "fc2.b"], resid=x, aux=dl)
                else:
                    ops.gemm(xn, P[k + "out.w"], x, mode=ops.EPI_BIAS_RESID_F32, bias=P[k + "out.b"], resid=x)
                    # ---- x = x + fc2(QuickGELU(fc1(norm2(x))))                                    utils.py:199
                    ops.layernorm(x, P[k + "norm2.w"], P[k + "norm2.b"], xn, rows=M, in_row_stride=D)
                    ops.gemm(xn, P[k + "fc1.w"], h, mode=ops.EPI_BIAS_GELU_BF16, bias=P[k + "fc1.b"])
                    ops.gemm(h, P[k + "fc2.w"], x, mode=ops.EPI_BIAS_RESID_F32, bias=P[k + "fc2.b"], resid=x)
            else:
                # Last layer: only the cls row of each frame reaches the output (vision_encoder.py:126), so every row-wise
                # op after the key/value projection runs on the BT cls rows alone: K/V for all rows, Q / attention query /
                # out-proj / MLP for one row per frame.  Row-independent ops => bit-for-meaning identical results.
                xn_cls = xn.view(BT, N0 * D)[:, :D]                      # row f*N0 of a [M, D] buffer, pitch N0*D
                x_cls_src = x.view(BT, N0 * D)[:, :D]
                if folded_in:
                    ops.gemm(W["xb"], P[k + "qkv.wf"][D:], qkv[:, D:], mode=ops.EPI_LNFOLD_BF16, bias=P[k + "qkv.bf"][D:],
                             stats_in=W["xstats"], eps=LN_EPS,
                             max_ctas=self.qkv_ctas if side is not None else 0)
                    ops.gemm(W["xb"].view(BT, N0 * D)[:, :D], P[k + "qkv.wf"][:D], qkv.view(BT, N0 * 3 * D)[:, :D],
                             mode=ops.EPI_LNFOLD_BF16, bias=P[k + "qkv.bf"][:D], stats_in=W["xstats"], stats_row_step=N0,
                             eps=LN_EPS)
                else:
                    ops.gemm(xn, P[k + "qkv.w"][D:], qkv[:, D:], mode=ops.EPI_BIAS_BF16, bias=P[k + "qkv.b"][D:],
                             max_ctas=self.qkv_ctas if side is not None else 0)
                    ops.gemm(xn_cls, P[k + "qkv.w"][:D], qkv.view(BT, N0 * 3 * D)[:, :D], mode=ops.EPI_BIAS_BF16,
                             bias=P[k + "qkv.b"][:D])
                if side is not None:
                    main.wait_event(ev_kv)
                ops.attention(qkv, W["kvaux"], xn, B=B, T=T, N0=N0, G=G, H=H, n_q=1, key_groups=kg)
                ops.gemm(xn_cls, P[k + "out.w"], W["x_cls"], mode=ops.EPI_BIAS_RESID_F32, bias=P[k + "out.b"],
                         resid=x_cls_src)
                ops.layernorm(W["x_cls"], P[k + "norm2.w"], P[k + "norm2.b"], W["cls_bf"], rows=BT, in_row_stride=D)
                ops.gemm(W["cls_bf"], P[k + "fc1.w"], h[:BT], mode=ops.EPI_BIAS_GELU_BF16, bias=P[k + "fc1.b"])
                ops.gemm(h[:BT], P[k + "fc2.w"], W["x_cls"], mode=ops.EPI_BIAS_RESID_F32, bias=P[k + "fc2.b"],
                         resid=W["x_cls"])
            if self.taps is not None and not torch.cuda.is_current_stream_capturing():
                self.taps.setdefault("layer_cls", []).append((x.view(BT, N0 * D)[:, :D] if i < L - 1 else W["x_cls"]).clone())
                self.taps.setdefault("layer_summary", []).append(W["summary"].clone())

        # ln_post(cls) @ proj per frame                                   vision_encoder.py:126-127
        ops.layernorm(W["x_cls"], P["ln_post.w"], P["ln_post.b"], W["cls_bf"], rows=BT, in_row_stride=D)
        ops.gemm(W["cls_bf"], P["proj.wT"], W["frame_feat"], mode=ops.EPI_BIAS_F32)
        return W["frame_feat"], (W["summary"] if e.use_summary_token else None)


    # ------------------------------------------------------------------ fp32-accuracy schedule
    def _ws_precise(self, B: int, T: int) -> Dict[str, Tensor]:
        key = ("precise", B, T)
        if key in self.ws:
            return self.ws[key]
        e, dev = self.enc, self.device
        D, N0, G = e.feature_dim, int(e.num_patches), e.num_global_prompts
        BT, M = B * T, B * T * int(e.num_patches)
        Dm = e.blocks[0].mlp.fc1.out_features
        kp = _ceil_to(3 * e.patch_embed.patch_size ** 2, 64)
        wide = max(Dm, kp)
        bf, f32 = torch.bfloat16, torch.float32
        z = lambda *s, dt: torch.empty(*s, device=dev, dtype=dt)
        W = {"hi": z(M * wide, dt=bf), "lo": z(M * wide, dt=bf),            # operand split scratch (largest: [M, 4D])
             "x": z(M, D, dt=f32), "t": z(M, D, dt=f32), "qkv": z(M, 3 * D, dt=f32), "att": z(M, D, dt=f32),
             "h": z(M, Dm, dt=f32), "ptmp": z(BT * (N0 - 1), D, dt=f32),
             "cls": z(BT, D, dt=f32), "cls_proj": z(BT, D, dt=f32), "sum_n": z(BT, D, dt=f32), "sqkv": z(BT, 3 * D, dt=f32),
             "smix": z(BT, D, dt=f32), "summary": z(BT, D, dt=f32), "aux_n": z(G + 2 * BT, D, dt=f32),
             "kvaux": z(G + 2 * BT, 2 * D, dt=f32), "x_cls": z(BT, D, dt=f32), "hc": z(BT, Dm, dt=f32),
             "frame_feat": z(BT, self.packed["proj.wT"].shape[0], dt=f32)}
        if not e.use_local_prompts:
            W["zero_local"] = torch.zeros(1, T, D, device=dev, dtype=f32)
        if not (e.use_summary_token or e.use_local_prompts):
            W["cls_proj"].zero_()
        if not e.use_summary_token:
            W["summary"].zero_()
        self.ws[key] = W
        return W

    def encode_frames_precise(self, video: Tensor) -> Tuple[Tensor, Tensor]:
        """Same schedule as encode_frames with fp32 activations; every contraction is a 3-term split-bf16 product."""
        e, P = self.enc, self.packed
        Q = self.pack_precise()
        B, T, _, _ = _clip_dims(video)
        D, N0, G, H = e.feature_dim, int(e.num_patches), e.num_global_prompts, e.num_heads
        BT, M = B * T, B * T * N0
        W = self._ws_precise(B, T)
        te = self.time_embed_for(T)
        x, t, qkv, att, h = W["x"], W["t"], W["qkv"], W["att"], W["h"]

        def mm(a32: Tensor, rows: int, K: int, wname, out: Tensor, bias=None, resid=None, w_rows=None, in_row_stride=None):
            """out = [resid +] a32[:rows, :K] @ w.T + bias, a32 fp32 (row pitch in_row_stride)."""
            hi = W["hi"][: rows * K].view(rows, K)
            lo = W["lo"][: rows * K].view(rows, K)
            ops.split_bf16(a32, hi, lo, rows=rows, D=K, in_row_stride=in_row_stride)
            w_hi, w_lo = Q[wname]
            if w_rows is not None:
                w_hi, w_lo = w_hi[w_rows], w_lo[w_rows]
            ops.gemm3(hi, lo, w_hi, w_lo, out, bias=bias, resid=resid)

        # patch embedding, cls rows, ln_pre
        kp = Q["patch.w"][0].shape[1]
        hi = W["hi"][: BT * (N0 - 1) * kp].view(-1, kp)
        lo = W["lo"][: BT * (N0 - 1) * kp].view(-1, kp)
        if video.dtype == torch.uint8:
            ops.patch_gather_u8(video, hi, lo, e.input_norm[0], e.input_norm[1], e.patch_embed.patch_size)
        else:
            ops.patch_gather_split(video, hi, lo, e.patch_embed.patch_size)
        ops.gemm3(hi, lo, *Q["patch.w"], W["ptmp"], bias=P["patch.b"])
        ops.embed_assemble(W["ptmp"], x, P["pos_embed"], te, BT, T, N0 - 1)
        ops.cls_rows(x, N0, P["cls_token"], P["pos_embed"], te, BT, T)
        ops.layernorm(x, P["ln_pre.w"], P["ln_pre.b"], x, rows=M, in_row_stride=D)

        L = len(e.blocks)
        kg = (1 if e.use_local_prompts else 0) | (2 if e.use_summary_token else 0)
        for i in range(L):
            k = f"b{i}."
            # temporal message (utils.py:160-188)
            if e.use_summary_token or e.use_local_prompts:
                mm(x, BT, D, k + "cls.w", W["cls_proj"], bias=P[k + "cls.b"], in_row_stride=N0 * D)
            if e.use_summary_token:
                ops.layernorm(W["cls_proj"], P[k + "summary_ln.w"], P[k + "summary_ln.b"], W["sum_n"], rows=BT, in_row_stride=D)
                mm(W["sum_n"], BT, D, k + "sqkv.w", W["sqkv"], bias=P[k + "sqkv.b"])
                ops.attention_f32(W["sqkv"], None, W["smix"], B=B, T=1, N0=T, G=0, H=H)
                mm(W["smix"], BT, D, k + "sout.w", W["summary"], bias=P[k + "sout.b"], resid=W["cls_proj"])
            ops.prompt_rows_f32(P["global_prompts"][i], P[k + "local"] if e.use_local_prompts else W["zero_local"],
                                W["cls_proj"], W["summary"], P[k + "norm1.w"], P[k + "norm1.b"], W["aux_n"], B, T, LN_EPS)
            mm(W["aux_n"], G + 2 * BT, D, k + "qkv.w", W["kvaux"], bias=P[k + "qkv.b"][D:], w_rows=slice(D, 3 * D))
            # x = x + attn(norm1(x))
            ops.layernorm(x, P[k + "norm1.w"], P[k + "norm1.b"], t, rows=M, in_row_stride=D)
            mm(t, M, D, k + "qkv.w", qkv, bias=P[k + "qkv.b"])
            ops.attention_f32(qkv, W["kvaux"], att, B=B, T=T, N0=N0, G=G, H=H, key_groups=kg)
            mm(att, M, D, k + "out.w", x, bias=P[k + "out.b"], resid=x)
            # x = x + fc2(QuickGELU(fc1(norm2(x))))
            ops.layernorm(x, P[k + "norm2.w"], P[k + "norm2.b"], t, rows=M, in_row_stride=D)
            mm(t, M, D, k + "fc1.w", h, bias=P[k + "fc1.b"])
            ops.quickgelu_f32(h)
            mm(h, M, h.shape[1], k + "fc2.w", x, bias=P[k + "fc2.b"], resid=x)

        ops.layernorm(x, P["ln_post.w"], P["ln_post.b"], W["cls"], rows=BT, in_row_stride=N0 * D)
        mm(W["cls"], BT, D, "proj.wT", W["frame_feat"])
        return W["frame_feat"], (W["summary"] if e.use_summary_token else None)


class CLIPVisionEncoder(nn.Module):
    """Same constructor / parameters / outputs as the reference class (vision_encoder.py:18-132).

    `precision` (attribute, default $GAVA_PRECISION or "bf16"): "bf16" = tensor-core bf16 operands, fp32 accumulation and
    residual stream (logits within 2e-2 of the fp32 reference); "fp32" = split-bf16 3-term products + fp32 glue (within
    1e-3; roughly 4x slower, a parity instrument)."""

    def __init__(self, input_size: Tuple[int, int] = (224, 224), num_frames: int = 8, feature_dim: int = 768,
                 patch_size: Tuple[int, int] = (16, 16), num_heads: int = 12, num_layers: int = 12,
                 mlp_factor: float = 4.0, act=QuickGELU, embed_dim: int = 512, use_summary_token: bool = False,
                 use_local_prompts: bool = False, use_global_prompts: bool = False, num_global_prompts: int = 8):
        super().__init__()
        self.feature_dim = feature_dim
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.num_frames = num_frames
        self.patch_embed = ImagePatchEmbed2D(img_size=input_size[0], patch_size=patch_size[0], in_chans=3,
                                             embed_dim=feature_dim)
        self.num_patches = (input_size[0] // patch_size[0]) * (input_size[1] // patch_size[1]) + 1  # incl. cls
        self.cls_token = nn.Parameter(torch.zeros([feature_dim]))
        self.pos_embed = nn.Parameter(torch.zeros([self.num_patches, feature_dim]))
        self.time_embed = nn.Parameter(torch.zeros([num_frames, feature_dim]))
        self.blocks = nn.ModuleList([
            TransformerEncoderLayer(in_feature_dim=feature_dim, qkv_dim=feature_dim, num_heads=num_heads,
                                    mlp_factor=mlp_factor, act=act, use_summary_token=use_summary_token,
                                    use_local_prompts=use_local_prompts, num_frames=num_frames, patch_size=patch_size)
            for _ in range(num_layers)])
        self.ln_pre = LayerNorm(feature_dim)
        self.ln_post = LayerNorm(feature_dim)
        self.proj = nn.Parameter(feature_dim ** -0.5 * torch.randn(feature_dim, embed_dim))
        self.use_summary_token = use_summary_token
        self.use_local_prompts = use_local_prompts
        self.use_global_prompts = use_global_prompts
        self.num_global_prompts = num_global_prompts
        if use_global_prompts:
            self.global_prompts = nn.Parameter(torch.zeros(num_layers, num_global_prompts, feature_dim))
            bound = _prompt_init_bound(patch_size, feature_dim)
            nn.init.uniform_(self.global_prompts.data, -bound, bound)
        for p in (self.cls_token, self.pos_embed, self.time_embed):
            nn.init.normal_(p, std=0.02)
        self._engine: Optional[TowerEngine] = None
        self.precision = os.environ.get("GAVA_PRECISION", "bf16")
        # (mean[3], std[3]) of the input stage; when set, forward() also accepts the decoder's uint8 (B,T,H,W,3) clips
        # and normalises them on the fly (video_dataset/dataset.py:117-122; dataloader defaults 0.45 / 0.225)
        self.input_norm = None

    def set_input_norm(self, mean, std) -> None:
        mean, std = [float(v) for v in mean], [float(v) for v in std]
        if len(mean) == 1:
            mean, std = mean * 3, std * 3
        assert len(mean) == 3 and len(std) == 3 and all(v > 0 for v in std)
        if self.input_norm != (mean, std) and self._engine is not None:
            # mean/std are passed BY VALUE to the uint8 patch gather and therefore baked into captured graphs
            for k in [k for k, v in self._engine.graphs.items() if v[1].dtype == torch.uint8]:
                del self._engine.graphs[k]
        self.input_norm = (mean, std)

    # -- engine plumbing ---------------------------------------------------------------------------
    def engine(self, device: torch.device) -> TowerEngine:
        if not self.use_global_prompts:
            # the reference encoder itself crashes without global prompts (VitaCLIP_vision_encoder.py:123-124 assigns the
            # layer's (x, summary) tuple to x): there is no behaviour to reproduce
            raise NotImplementedError("use_global_prompts=False: the reference encoder has no working forward for it")
        if self.feature_dim // self.num_heads != 64:
            raise NotImplementedError("attention kernel is specialised for head_dim 64")
        if self._engine is None or self._engine.device != device:
            from .train_engine import TowerTrainer           # TowerEngine + tape / backward schedule
            self._engine = TowerTrainer(self, device)
        self._engine.pack()
        return self._engine

    def _check_input(self, x: Tensor) -> Tensor:
        u8 = x.dtype == torch.uint8
        if u8:
            if x.dim() != 5 or x.shape[-1] != 3:
                raise ValueError(f"expected uint8 clip of shape (B,T,H,W,3), got {tuple(x.shape)}")
            if self.input_norm is None:
                raise RuntimeError("uint8 clips need set_input_norm(mean, std) first")
        elif x.dim() != 5 or x.shape[1] != 3:
            raise ValueError(f"expected video of shape (B,3,T,H,W), got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("gava_clip_b200 runs on sm_100a only: move the clip to a CUDA device "
                               "(there is no CPU fallback)")
        if x.shape[1 if u8 else 2] != self.num_frames:
            # the reference blocks hard-code num_frames (utils.py:161): other lengths mis-shape its views
            raise ValueError(f"clip has T={x.shape[1 if u8 else 2]} frames but the encoder was built with num_frames={self.num_frames}")
        return x.contiguous() if u8 else x.contiguous().float()

    def _wants_grad(self) -> bool:
        """training mode + grad enabled + a trainable tower parameter: the forward keeps a tape and is differentiable."""
        return self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def frames_forward(self, x: Tensor) -> Tuple[Tensor, Tensor, int, int]:
        """per-frame projected cls features [BT, E(+pad)] and summary tokens [BT, D] (workspace views)."""
        x = self._check_input(x)
        eng = self.engine(x.device)
        B, T, _, _ = _clip_dims(x)
        if self._wants_grad():
            if self.precision == "fp32":
                raise NotImplementedError("the fine-tune step runs in the bf16 path (set precision='bf16')")
            from .train_engine import TowerFunction
            eng.check_trainable()
            ff, st = TowerFunction.apply(eng, x, self.embed_dim, *[p for _, p in eng.trainable_named()])
            return ff, st, B, T
        ff, st = eng.run(x)
        return ff, st, B, T

    def forward(self, x: Tensor):
        if x.is_cuda and x.device.index != torch.cuda.current_device():
            with torch.cuda.device(x.device):        # the kernels launch on the current device / stream
                return self.forward(x)
        ff, st, B, T = self.frames_forward(x)
        E = self.embed_dim
        # temporal means (vision_encoder.py:128-130); the fused head kernel does the same inside VitaCLIP.forward
        cls_x = ff[:, :E].reshape(B, T, E).mean(dim=1)
        summary = st.reshape(B, T, -1).mean(dim=1) if st is not None else None      # vision_encoder.py:129-130
        return cls_x, summary
