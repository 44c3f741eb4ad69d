"""Fine-tune step of the video tower on the sm_100a kernels (reference: training/train.py:441-494).

The backbone is frozen (VitaCLIP_model.py:230-235): gradients flow THROUGH it (input gradients only, "dgrad") and are
collected for the parameters the reference leaves trainable -- names containing `summary`, `local`, `global`,
`time_embed`: visual.time_embed, visual.global_prompts, blocks.{i}.local_prompts, blocks.{i}.summary_ln.*,
blocks.{i}.summary_attn_layer.{q,k,v,out}_proj.*.

`TowerTrainer` extends the inference `TowerEngine` (vision_encoder.py):
  * forward_train(video): the same kernel schedule with every layer's activations kept on a tape in HBM (x_in, x_mid fp32;
    qkv, kv_aux, attention output, fc1 pre-activations bf16; the small temporal-message tensors);
  * backward(d_frame_feat, d_summary): dgrad GEMMs on pre-transposed frozen weights (tcgen05 GEMM, DGELU epilogue through
    the MLP), flash-style attention backward, LayerNorm backward kernels, and for the trainable tensors the small wgrad
    GEMMs (dY^T X on transposed operands), row reductions and the prompt-row scatter.  All gradients land in ONE flat
    fp32 buffer (`grad_flat`), the unit of the NCCL all-reduce and of the fused AdamW kernel (finetune.py).
`TowerFunction` exposes the pair to torch.autograd so that the drop-in VitaCLIP.forward -> loss.backward() works unchanged
(also under DistributedDataParallel, whose hooks fire on the parameters' .grad accumulation).

There is no PyTorch fallback: every arithmetic op of the tower's forward and backward is a kernel behind the C ABI.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .vision_encoder import LN_EPS, TowerEngine, _ceil_to, _clip_dims

Tensor = torch.Tensor
TRAINABLE_KEYS = ("summary", "local", "global", "time_embed")      # VitaCLIP_model.py:230-235


def is_trainable_name(name: str) -> bool:
    return any(k in name for k in TRAINABLE_KEYS)


class TowerTrainer(TowerEngine):
    # ------------------------------------------------------------------ parameters / gradients
    def trainable_named(self) -> List[Tuple[str, torch.nn.Parameter]]:
        """(name, parameter) of the tensors this build can produce gradients for, in named_parameters() order."""
        return [(n, p) for n, p in self.enc.named_parameters() if is_trainable_name(n)]

    def check_trainable(self) -> None:
        bad = [n for n, p in self.enc.named_parameters() if p.requires_grad and not is_trainable_name(n)]
        if bad:
            raise NotImplementedError(
                "the sm_100a backward covers the reference's fine-tune recipe (frozen backbone; trainable: names containing "
                f"{TRAINABLE_KEYS}); these parameters also require grad: {bad[:4]}{'...' if len(bad) > 4 else ''}")

    def _alloc_grads(self) -> None:
        """One flat fp32 buffer; q/k/v of a summary attention layer are adjacent so one wgrad GEMM fills all three."""
        if getattr(self, "grad_flat", None) is not None:
            return
        e, dev = self.enc, self.device
        D, L, G, T = e.feature_dim, len(e.blocks), e.num_global_prompts, e.num_frames
        layout: List[Tuple[str, Tuple[int, ...]]] = [("time_embed", (T, D)), ("global_prompts", (L, G, D))]
        bounds: List[Tuple[int, int]] = []             # [start, end) of each layer's entries in `layout`
        for i in range(L):
            k = f"blocks.{i}."
            first = len(layout)
            if e.use_local_prompts:                    # utils.py:137-139
                layout += [(k + "local_prompts", (1, T, D))]
            if e.use_summary_token:                    # utils.py:129-134
                layout += [(k + "summary_ln.weight", (D,)), (k + "summary_ln.bias", (D,)),
                           (k + "sqkv.weight", (3 * D, D)), (k + "sqkv.bias", (3 * D,)),
                           (k + "summary_attn_layer.out_proj.weight", (D, D)), (k + "summary_attn_layer.out_proj.bias", (D,))]
            bounds.append((first, len(layout)))
        sizes = [int(torch.Size(s).numel()) for _, s in layout]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        self.grad_flat = torch.zeros(max(offs[-1], 1), device=dev, dtype=torch.float32)
        self.grads: Dict[str, Tensor] = {}
        for (name, shape), off, n in zip(layout, offs, sizes):
            self.grads[name] = self.grad_flat[off: off + n].view(shape)
        # contiguous per-layer ranges of grad_flat (all-reduce units; empty when the layer has no trainable tensor)
        self.layer_slices: List[Tensor] = [self.grad_flat[offs[a]: offs[b]] for a, b in bounds]
        self.head_slice = self.grad_flat[: T * D + L * G * D]       # time_embed + global_prompts
        if e.use_summary_token:
            for i in range(L):       # per-parameter views of the fused q|k|v blocks
                k = f"blocks.{i}."
                w, b = self.grads[k + "sqkv.weight"], self.grads[k + "sqkv.bias"]
                for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                    self.grads[k + f"summary_attn_layer.{nm}.weight"] = w[j * D:(j + 1) * D]
                    self.grads[k + f"summary_attn_layer.{nm}.bias"] = b[j * D:(j + 1) * D]

    def grad_of(self, name: str) -> Tensor:
        return self.grads[name]

    # ------------------------------------------------------------------ transposed frozen weights (dgrad operands)
    def pack_train(self) -> None:
        """W^T copies: dX[M,K] = dY[M,N] W[N,K] runs as the forward GEMM with W^T [K,N] as its [N', K'] operand."""
        self.pack()
        e, P = self.enc, self.packed
        vers = self.__dict__.setdefault("_t_versions", {})

        def put_t(name: str, src: Tensor) -> None:
            key = (src.data_ptr(), src._version)
            if name in P and vers.get(name) == key:
                return
            if name in P and P[name].shape == (src.shape[1], src.shape[0]):
                ops.transpose_bf16(src.detach(), P[name])   # bf16 [N,K] -> [K,N]
            else:
                P[name] = src.detach().t().to(self.device, torch.bfloat16).contiguous()
                self._t_moves = getattr(self, "_t_moves", 0) + 1
            vers[name] = key

        D = e.feature_dim
        if "proj.w" not in P or vers.get("proj.w") != (e.proj.data_ptr(), e.proj._version):
            pw = torch.zeros(D, _ceil_to(e.embed_dim, 128), device=self.device, dtype=torch.bfloat16)
            pw[:, : e.embed_dim] = e.proj.detach().to(self.device, torch.bfloat16)
            P["proj.w"] = pw
            self._t_moves = getattr(self, "_t_moves", 0) + 1
            vers["proj.w"] = (e.proj.data_ptr(), e.proj._version)
        for i in range(len(e.blocks)):
            k = f"b{i}."
            for nm in ("qkv", "out", "fc1", "fc2", "cls", "sqkv", "sout"):
                if k + nm + ".w" in P:          # cls / sqkv / sout exist only with the prompt groups that use them
                    put_t(k + nm + ".wT", P[k + nm + ".w"])

    # ------------------------------------------------------------------ tape + scratch
    def train_ws(self, B: int, T: int) -> Dict[str, object]:
        key = ("train", B, T)
        if key in self.ws:
            return self.ws[key]
        # a tape is ~1.9 GB per layer at config 2: keep at most two clip shapes alive (the full batch and a ragged tail)
        for old in [k for k in self.ws if isinstance(k, tuple) and k and k[0] == "train"][:-1]:
            del self.ws[old]
        e, dev = self.enc, self.device
        D, N0, G, L = e.feature_dim, int(e.num_patches), e.num_global_prompts, len(e.blocks)
        BT, M = B * T, B * T * N0
        Dm = e.blocks[0].mlp.fc1.out_features
        R = G + 2 * BT
        BTp = _ceil_to(BT, 64)
        Ep = self.packed["proj.wT"].shape[0]
        bf, f32 = torch.bfloat16, torch.float32
        z = lambda *s, dt: torch.empty(*s, device=dev, dtype=dt)
        per_layer = lambda *s, dt: [z(*s, dt=dt) for _ in range(L)]
        W: Dict[str, object] = {
            # ---- tape
            "x0_pre": z(M, D, dt=f32), "x_in": per_layer(M, D, dt=f32), "x_mid": [z(M, D, dt=f32) for _ in range(L - 1)],
            "qkv": per_layer(M, 3 * D, dt=bf), "kvaux": per_layer(R, 2 * D, dt=bf), "o": per_layer(M, D, dt=bf),
            "lse": per_layer(BT * e.num_heads * N0, dt=f32),
            "a1": [z(M, Dm, dt=bf) for _ in range(L - 1)], "cls_proj": per_layer(BT, D, dt=f32), "summary": per_layer(BT, D, dt=f32),
            "sum_n": per_layer(BT, D, dt=bf), "sqkv": per_layer(BT, 3 * D, dt=bf), "smix": per_layer(BT, D, dt=bf),
            # ---- forward scratch
            "patches": z(BT * (N0 - 1), self.packed["patch.w"].shape[1], dt=bf), "xn": z(M, D, dt=bf), "h": z(M, Dm, dt=bf),
            "cls_bf": z(BT, D, dt=bf), "cls_bf2": z(BT, D, dt=bf), "aux_n": z(R, D, dt=bf), "frame_feat": z(BT, Ep, dt=f32),
            # ---- backward scratch
            "dx": z(M, D, dt=f32), "dxb": z(M, D, dt=bf), "g4": z(M, Dm, dt=bf), "gD": z(M, D, dt=bf),
            "dqkv": z(M, 3 * D, dt=bf), "dkvaux": z(R, 2 * D, dt=f32),
            "dkv_slots": z(BT * e.num_heads * (G + T + 1) * 128, dt=f32), "dkvaux_bf": z(R, 2 * D, dt=bf),
            "daux_n": z(R, D, dt=bf), "d_lcp": z(BT, D, dt=f32), "d_sum": z(BT, D, dt=f32), "d_sum_bf": z(BT, D, dt=bf),
            "dsmix": z(BT, D, dt=bf), "dsqkv": z(BT, 3 * D, dt=bf), "dsn": z(BT, D, dt=bf), "d_cp": z(BT, D, dt=f32),
            "d_cp_bf": z(BT, D, dt=bf), "tA": z(3 * D, BTp, dt=bf), "tB": z(D, BTp, dt=bf),
            "dff_bf": z(BT, Ep, dt=bf), "dcls_n": z(BT, D, dt=bf), "dff": z(BT, Ep, dt=f32),
            # ---- last layer, cls rows only (vision_encoder.py:126: nothing else reaches the output)
            "xc_mid": z(BT, D, dt=f32), "xc_out": z(BT, D, dt=f32), "a1c": z(BT, Dm, dt=bf), "hc": z(BT, Dm, dt=bf),
            "dxc": z(BT, D, dt=f32), "dxc_bf": z(BT, D, dt=bf), "g4c": z(BT, Dm, dt=bf), "gDc": z(BT, D, dt=bf),
            "dxc2": z(BT, D, dt=f32), "dxc2_bf": z(BT, D, dt=bf), "d_sum_ext": z(BT, D, dt=f32),
        }
        W["gx0"] = W["g4"].view(-1)[: 2 * M * D].view(torch.float32).view(M, D)     # tail scratch: g4 is dead by then
        # inputs of the prompt-row kernels that a disabled prompt group never produces (finite values; the rows are masked)
        if not e.use_local_prompts:
            W["zero_local"] = torch.zeros(1, T, D, device=dev, dtype=f32)
        if not e.use_summary_token:
            for t in W["summary"]:
                t.zero_()
            if not e.use_local_prompts:
                for t in W["cls_proj"]:
                    t.zero_()
        self.ws[key] = W
        return W

    # ------------------------------------------------------------------ forward with tape
    def forward_train(self, video: Tensor) -> Tuple[Tensor, Tensor]:
        """Same schedule as TowerEngine.encode_frames (incl. the cls-rows-only last layer), activations kept for backward().

        The ~190 launches are replayed from one CUDA graph per clip shape (the tape, the packed weights and the static clip
        buffer keep their addresses; re-packing after an optimizer step copies in place): the eager fine-tune step spends
        ~24 ms of host time issuing launches, which bounds small-batch recipes (70 frames x 4 clips per GPU)."""
        self.pack_train()
        B, T, _, _ = _clip_dims(video)
        W = self.train_ws(B, T)
        # every forward stamps its tape: backward() refuses a tape that a later forward of the same shape overwrote
        self.tape_generation = getattr(self, "tape_generation", 0) + 1
        W["generation"] = self.tape_generation
        L = len(self.enc.blocks)
        if not self._train_graph_ok(T):
            n0 = ops.launch_count()
            self._forward_train_body(video, W)
            self.launches_last_call = ops.launch_count() - n0
            return W["frame_feat"], (W["summary"][L - 1] if self.enc.use_summary_token else None)
        key = ("fwd", tuple(video.shape), video.dtype)
        ent = self._train_graphs(W).get(key)
        if ent is None:
            static_in = torch.empty_like(video)
            static_in.copy_(video)
            self._forward_train_body(static_in, W)        # warm-up: kernel attributes, allocator
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):   # a DataLoader thread may be busy with CUDA calls
                self._forward_train_body(static_in, W)
            ent = (g, static_in, ops.launch_count() - n0)
            self._train_graphs(W)[key] = ent
        g, static_in, n_launch = ent
        if video.data_ptr() != static_in.data_ptr():
            static_in.copy_(video)
        g.replay()
        self.launches_last_call = n_launch
        return W["frame_feat"], (W["summary"][L - 1] if self.enc.use_summary_token else None)

    def train_input_buffer(self, shape, dtype=torch.float32) -> Optional[Tensor]:
        """Static clip buffer of the captured training forward for this shape (None before the first step): a caller that
        uploads straight into it skips the staging copy."""
        B, T = shape[0], shape[2]
        W = self.ws.get(("train", B, T))
        ent = W.get("graphs", {}).get(("fwd", tuple(shape), dtype)) if W is not None else None
        return ent[1] if ent is not None else None

    # ---- CUDA-graph bookkeeping of the training step
    def _train_graphs(self, W: Dict[str, object]) -> Dict[tuple, tuple]:
        """Graphs of the training forward / backward segments of one tape (they die with it); dropped when a packed
        operand moved (pack() / pack_train() re-allocated instead of copying in place)."""
        stamp = (self.pack_moves, getattr(self, "_t_moves", 0))
        if W.get("graph_stamp") != stamp:
            W["graphs"], W["graph_stamp"] = {}, stamp
        return W["graphs"]

    def _train_graph_ok(self, T: int) -> bool:
        # a resized time_embed copy is rebuilt (new address) whenever time_embed changes, i.e. after every optimizer step
        return (self.use_graph and not torch.cuda.is_current_stream_capturing()
                and self.packed["time_embed"].shape[0] == T)

    def _forward_train_body(self, video: Tensor, W: Dict[str, object]) -> None:
        e, P = self.enc, self.packed
        B, T, _, _ = _clip_dims(video)
        D, N0, G, H = e.feature_dim, int(e.num_patches), e.num_global_prompts, e.num_heads
        BT, M = B * T, B * T * N0
        te = self.time_embed_for(T)
        xn, h = W["xn"], W["h"]

        if video.dtype == torch.uint8:
            ops.patch_gather_u8(video, W["patches"], None, e.input_norm[0], e.input_norm[1], e.patch_embed.patch_size)
        else:
            ops.patch_gather(video, W["patches"], e.patch_embed.patch_size)
        x0 = W["x0_pre"]
        ops.gemm(W["patches"], P["patch.w"], x0, mode=ops.EPI_PATCH_F32, bias=P["patch.b"], pos=P["pos_embed"],
                 time_embed=te, group=N0 - 1, group_stride=N0, row_offset=1, T=T)
        ops.cls_rows(x0, N0, P["cls_token"], P["pos_embed"], te, BT, T)
        ops.layernorm2(x0, P["ln_pre.w"], P["ln_pre.b"], W["x_in"][0], P["b0.norm1.w"], P["b0.norm1.b"], xn, rows=M)

        L = len(e.blocks)
        main = torch.cuda.current_stream(self.device)
        side = self.side_stream if self.overlap_chain else None
        cap = self.qkv_ctas if side is not None else 0
        kg = (1 if e.use_local_prompts else 0) | (2 if e.use_summary_token else 0)      # prompt groups among the keys
        for i in range(L):
            k = f"b{i}."
            xin = W["x_in"][i]
            xmid, xnext = (W["x_mid"][i], W["x_in"][i + 1]) if i < L - 1 else (None, None)
            cp, sm = W["cls_proj"][i], W["summary"][i]
            # temporal message (utils.py:160-188): a side-stream branch next to norm1 + the QKV GEMM (whose grid is capped
            # so that these small launches find free SMs), joined before the attention kernel -- as in TowerEngine
            if side is not None:
                ev_x = torch.cuda.Event()
                ev_x.record(main)
                side.wait_event(ev_x)
            with torch.cuda.stream(side if side is not None else main):
                if e.use_summary_token or e.use_local_prompts:                               # utils.py:160-165
                    ops.layernorm(xin, None, None, W["cls_bf2"], rows=BT, in_row_stride=N0 * D)
                    ops.gemm(W["cls_bf2"], P[k + "cls.w"], cp, mode=ops.EPI_BIAS_F32, bias=P[k + "cls.b"])
                if e.use_summary_token:                                                      # utils.py:168-172
                    ops.layernorm(cp, P[k + "summary_ln.w"], P[k + "summary_ln.b"], W["sum_n"][i], rows=BT, in_row_stride=D)
                    ops.gemm(W["sum_n"][i], P[k + "sqkv.w"], W["sqkv"][i], mode=ops.EPI_BIAS_BF16, bias=P[k + "sqkv.b"])
                    ops.temporal_attention(W["sqkv"][i], W["smix"][i], B, T, H)
                    ops.gemm(W["smix"][i], P[k + "sout.w"], sm, mode=ops.EPI_BIAS_RESID_F32, bias=P[k + "sout.b"], resid=cp)
                # rows of a disabled prompt group are built from zeros and masked out by the attention kernels (key_groups)
                ops.prompt_rows(P["global_prompts"][i], P[k + "local"] if e.use_local_prompts else W["zero_local"], cp, sm,
                                P[k + "norm1.w"], P[k + "norm1.b"], W["aux_n"], B, T, LN_EPS)
                ops.gemm(W["aux_n"], P[k + "qkv.w"][D:], W["kvaux"][i], mode=ops.EPI_BIAS_BF16, bias=P[k + "qkv.b"][D:])
                if side is not None:
                    ev_kv = torch.cuda.Event()
                    ev_kv.record(side)
            # x = x + attn(norm1(x))   (utils.py:190-191); block 0's norm1 came out of the ln_pre pass
            if i > 0:
                ops.layernorm(xin, P[k + "norm1.w"], P[k + "norm1.b"], xn, rows=M, in_row_stride=D)
            if i == L - 1:
                # Last layer: only the cls row of each frame reaches the output, so -- exactly as in the inference
                # schedule -- K/V are projected for all rows, Q / attention query / out-proj / MLP for the cls rows alone.
                qkv = W["qkv"][i]
                xn_cls = xn.view(BT, N0 * D)[:, :D]
                ops.gemm(xn, P[k + "qkv.w"][D:], qkv[:, D:], mode=ops.EPI_BIAS_BF16, bias=P[k + "qkv.b"][D:], max_ctas=cap)
                ops.gemm(xn_cls, P[k + "qkv.w"][:D], qkv.view(BT, N0 * 3 * D)[:, :D], mode=ops.EPI_BIAS_BF16,
                         bias=P[k + "qkv.b"][:D])
                if side is not None:
                    main.wait_event(ev_kv)
                ops.attention_train(qkv, W["kvaux"][i], W["o"][i], W["lse"][i], B=B, T=T, N0=N0, G=G, H=H, n_q=1, key_groups=kg)
                ops.gemm(W["o"][i].view(BT, N0 * D)[:, :D], P[k + "out.w"], W["xc_mid"], mode=ops.EPI_BIAS_RESID_F32,
                         bias=P[k + "out.b"], resid=xin.view(BT, N0 * D)[:, :D])
                ops.layernorm(W["xc_mid"], P[k + "norm2.w"], P[k + "norm2.b"], W["cls_bf"], rows=BT, in_row_stride=D)
                ops.gemm(W["cls_bf"], P[k + "fc1.w"], W["hc"], mode=ops.EPI_BIAS_GELU_BF16, bias=P[k + "fc1.b"],
                         aux_out=W["a1c"])
                ops.gemm(W["hc"], P[k + "fc2.w"], W["xc_out"], mode=ops.EPI_BIAS_RESID_F32, bias=P[k + "fc2.b"],
                         resid=W["xc_mid"])
                break
            ops.gemm(xn, P[k + "qkv.w"], W["qkv"][i], mode=ops.EPI_BIAS_BF16, bias=P[k + "qkv.b"], max_ctas=cap)
            if side is not None:
                main.wait_event(ev_kv)
            ops.attention_train(W["qkv"][i], W["kvaux"][i], W["o"][i], W["lse"][i], B=B, T=T, N0=N0, G=G, H=H, key_groups=kg)
            ops.gemm(W["o"][i], P[k + "out.w"], xmid, mode=ops.EPI_BIAS_RESID_F32, bias=P[k + "out.b"], resid=xin)
            # x = x + fc2(QuickGELU(fc1(norm2(x))))   (utils.py:199); the pre-activations stay for the DGELU epilogue
            ops.layernorm(xmid, P[k + "norm2.w"], P[k + "norm2.b"], xn, rows=M, in_row_stride=D)
            ops.gemm(xn, P[k + "fc1.w"], h, mode=ops.EPI_BIAS_GELU_BF16, bias=P[k + "fc1.b"], aux_out=W["a1"][i])
            ops.gemm(h, P[k + "fc2.w"], xnext, mode=ops.EPI_BIAS_RESID_F32, bias=P[k + "fc2.b"], resid=xmid)

        ops.layernorm(W["xc_out"], P["ln_post.w"], P["ln_post.b"], W["cls_bf"], rows=BT, in_row_stride=D)
        ops.gemm(W["cls_bf"], P["proj.wT"], W["frame_feat"], mode=ops.EPI_BIAS_F32)

    # ------------------------------------------------------------------ backward
    def backward(self, B: int, T: int, d_frame_feat: Optional[Tensor], d_summary: Optional[Tensor],
                 generation: Optional[int] = None) -> Dict[str, Tensor]:
        """Gradients of the trainable tower parameters given d(frame_feat) [BT, E(+pad)] and d(summary tokens) [BT, D]
        (either may be None = zero).  Returns self.grads (views into self.grad_flat); uses the tape of the last
        forward_train call for this (B, T) -- `generation` (the stamp that forward returned) guards against a tape that
        was evicted or overwritten by a later forward in between."""
        e, P = self.enc, self.packed
        tape = self.ws.get(("train", B, T))
        if tape is None:
            raise RuntimeError(f"backward: the activation tape of the ({B},{T}) forward was evicted (more than two clip "
                               "shapes in flight); run forward and backward of one batch back to back")
        if generation is not None and tape.get("generation") != generation:
            raise RuntimeError("backward: the activation tape was overwritten by a later train-mode forward of the same clip "
                               "shape (two forwards before one backward are not supported: the tape is per shape)")
        self._alloc_grads()
        L = len(e.blocks)
        W = self.train_ws(B, T)
        BT = B * T
        hook = getattr(self, "grad_ready_hook", None)   # finetune.FineTuner: per-layer gradient all-reduce overlap
        pending: List = []
        n0 = ops.launch_count()

        # ---- dynamic inputs -> static buffers (outside the replayed segments)
        have_ff, have_ext = d_frame_feat is not None, d_summary is not None
        if have_ff:
            dff = W["dff"]
            if d_frame_feat.shape[1] != dff.shape[1]:
                dff.zero_()
                dff[:, : d_frame_feat.shape[1]].copy_(d_frame_feat)
            else:
                dff.copy_(d_frame_feat)
        if have_ext:
            W["d_sum_ext"].copy_(d_summary.reshape(BT, -1))

        # ---- segments: [head + last layer], one per remaining layer, [tail]; without an all-reduce hook they fuse into one
        use_graph = self._train_graph_ok(T)
        graphs = self._train_graphs(W) if use_graph else None

        def run_segment(name: str, body) -> None:
            """A segment consumes the gradient stream in place (dx += ...), so it must execute exactly once per step: the
            first step runs it eagerly (kernel attributes, allocator), the second captures it (nothing executes during a
            capture) and replays, every later step only replays."""
            if graphs is None:
                body()
                return
            key = ("bwd", name, have_ff, have_ext)
            ent = graphs.get(key)
            if ent is None:
                graphs[key] = "seen"
                body()
                return
            if ent == "seen":
                g = torch.cuda.CUDAGraph()
                m0 = ops.launch_count()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):   # autograd runs this on its own thread
                    body()
                ent = (g, ops.launch_count() - m0)
                graphs[key] = ent
                n_warm[0] += ent[1]
            ent[0].replay()
            n_replayed[0] += ent[1]

        n_warm, n_replayed = [0], [0]
        if hook is None:
            def whole():
                self._bwd_head(W, B, T, have_ff)
                for i in reversed(range(L)):
                    self._bwd_layer(W, B, T, i, have_ext)
                self._bwd_tail(W, B, T)
            run_segment("all", whole)
        else:
            for i in reversed(range(L)):
                def seg(i=i):
                    if i == L - 1:
                        self._bwd_head(W, B, T, have_ff)
                    self._bwd_layer(W, B, T, i, have_ext)
                run_segment(f"layer{i}", seg)
                if self.layer_slices[i].numel():
                    pending.append(hook(self.layer_slices[i]))   # this layer's gradients are final: start their all-reduce now
            run_segment("tail", lambda: self._bwd_tail(W, B, T))
            pending.append(hook(self.head_slice))
            for h in pending:           # stream-level wait: autograd adds these buffers into the parameters' .grad next
                if h is not None:
                    h.wait()
        # kernels executed by this call: eager launches (minus the ones only recorded during a capture) + replayed ones
        self.bwd_launches_last_call = ops.launch_count() - n0 - n_warm[0] + n_replayed[0]
        return self.grads

    def _bwd_head(self, W: Dict[str, object], B: int, T: int, have_ff: bool) -> None:
        """frame_feat = ln_post(x_cls) @ proj   (vision_encoder.py:126-127): d(frame_feat) in W["dff"] -> W["dxc"]."""
        e, P = self.enc, self.packed
        D, BT = e.feature_dim, B * T
        W["dx"].zero_()
        if have_ff:
            Ep = W["dff"].shape[1]
            ops.layernorm(W["dff"], None, None, W["dff_bf"], rows=BT, in_row_stride=Ep)            # f32 -> bf16 cast
            ops.gemm(W["dff_bf"], P["proj.w"], W["dcls_n"], mode=ops.EPI_BIAS_BF16)
            ops.layernorm_bwd(W["dcls_n"], W["xc_out"], P["ln_post.w"], W["dxc"], rows=BT, D=D, dx_bf16=W["dxc_bf"])
        else:
            W["dxc"].zero_(); W["dxc_bf"].zero_()

    def _bwd_layer(self, W: Dict[str, object], B: int, T: int, i: int, have_ext: bool) -> None:
        e, P, g = self.enc, self.packed, self.grads
        D, N0, G, H = e.feature_dim, int(e.num_patches), e.num_global_prompts, e.num_heads
        BT, M = B * T, B * T * N0
        R = G + 2 * BT
        L = len(e.blocks)
        dx, dxb, gD = W["dx"], W["dxb"], W["gD"]
        dx_cls = dx.view(BT, N0 * D)[:, :D]
        ext = W["d_sum_ext"] if have_ext else None
        kg = (1 if e.use_local_prompts else 0) | (2 if e.use_summary_token else 0)
        k, kb = f"b{i}.", f"blocks.{i}."
        cp, sm = W["cls_proj"][i], W["summary"][i]
        if i == L - 1:
            # ---- last layer: MLP / out-proj / attention-query gradients exist for the BT cls rows only
            ops.gemm(W["dxc_bf"], P[k + "fc2.wT"], W["g4c"], mode=ops.EPI_DGELU_BF16, aux=W["a1c"])
            ops.gemm(W["g4c"], P[k + "fc1.wT"], W["gDc"], mode=ops.EPI_BIAS_BF16)
            # dxc itself stays untouched (the segment may be replayed): the sum lands in the stream's cls rows
            ops.layernorm_bwd(W["gDc"], W["xc_mid"], P[k + "norm2.w"], W["dxc2"], rows=BT, D=D, add1=W["dxc"],
                              dx_bf16=W["dxc2_bf"])
            dx_cls.copy_(W["dxc2"])                                              # residual path into the stream's cls rows
            ops.gemm(W["dxc2_bf"], P[k + "out.wT"], gD.view(BT, N0 * D)[:, :D], mode=ops.EPI_BIAS_BF16)   # dO, cls rows of gD
            W["dqkv"].view(BT * N0, 3, D)[:, 0].zero_()                          # dq of the rows without a query
            ops.attention_bwd(W["qkv"][i], W["kvaux"][i], W["o"][i], gD, W["dqkv"], W["dkvaux"], B=B, T=T, N0=N0, G=G,
                              H=H, lse=W["lse"][i], n_q=1, scratch=W["dkv_slots"], key_groups=kg)
        else:
            # ---- MLP: dx += dLN2( (dx W2 * g'(a1)) W1 )
            ops.gemm(dxb, P[k + "fc2.wT"], W["g4"], mode=ops.EPI_DGELU_BF16, aux=W["a1"][i])
            ops.gemm(W["g4"], P[k + "fc1.wT"], gD, mode=ops.EPI_BIAS_BF16)
            ops.layernorm_bwd(gD, W["x_mid"][i], P[k + "norm2.w"], dx, rows=M, D=D, add1=dx, dx_bf16=dxb)
            # ---- attention: dO = dx Wo; (dq, dk, dv, d kv_aux); dx += dLN1(dqkv Wqkv)
            ops.gemm(dxb, P[k + "out.wT"], gD, mode=ops.EPI_BIAS_BF16)
            ops.attention_bwd(W["qkv"][i], W["kvaux"][i], W["o"][i], gD, W["dqkv"], W["dkvaux"], B=B, T=T, N0=N0, G=G,
                              H=H, lse=W["lse"][i], scratch=W["dkv_slots"], key_groups=kg)
        # The K/V-only prompt rows' backward (~20 latency-bound launches on G + 2 BT rows, everything from d kv_aux down to
        # d cls_proj and the layer's parameter gradients) depends only on the attention backward's d kv_aux and touches none
        # of the stream's buffers: it runs on the side stream next to the qkv^T dgrad GEMM (grid capped, as in the forward)
        # and the LayerNorm backward, and joins before its result is added into the cls rows of dx.
        main = torch.cuda.current_stream(self.device)
        side = self.side_stream if self.overlap_chain else None
        if side is not None:
            ev_a = torch.cuda.Event()
            ev_a.record(main)
            side.wait_event(ev_a)
        ops.gemm(W["dqkv"], P[k + "qkv.wT"], gD, mode=ops.EPI_BIAS_BF16, max_ctas=self.qkv_ctas if side is not None else 0)
        ops.layernorm_bwd(gD, W["x_in"][i], P[k + "norm1.w"], dx, rows=M, D=D, add1=dx, dx_bf16=dxb)
        with torch.cuda.stream(side if side is not None else main):
            # ---- K/V-only prompt rows: d kv_aux -> normalised rows -> global / local / summary / cls_proj gradients
            ops.layernorm(W["dkvaux"].view(2 * R, D), None, None, W["dkvaux_bf"].view(2 * R, D), rows=2 * R, in_row_stride=D)
            ops.gemm(W["dkvaux_bf"], P[k + "qkv.wT"][:, D:], W["daux_n"], mode=ops.EPI_BIAS_BF16)
            # (rows of a disabled prompt group had zero probability: their d kv_aux, hence d_lcp / d_sum, are exact zeros)
            ops.prompt_rows_bwd(W["daux_n"], P["global_prompts"][i], P[k + "local"] if e.use_local_prompts else W["zero_local"],
                                cp, sm, P[k + "norm1.w"], g["global_prompts"][i], W["d_lcp"], W["d_sum"],
                                ext if (i == L - 1 and e.use_summary_token) else None, B, T, LN_EPS)
            if e.use_local_prompts:
                ops.group_sum(W["d_lcp"], g[kb + "local_prompts"], outer=B, J=T, inner=1)
            if e.use_summary_token:
                # ---- summary = cls_proj + out_proj(temporal_attention(qkv(summary_ln(cls_proj))))   (utils.py:168-170)
                ops.layernorm(W["d_sum"], None, None, W["d_sum_bf"], rows=BT, in_row_stride=D)     # cast
                ops.gemm(W["d_sum_bf"], P[k + "sout.wT"], W["dsmix"], mode=ops.EPI_BIAS_BF16)
                self._wgrad(W, W["d_sum"], W["smix"][i], g[kb + "summary_attn_layer.out_proj.weight"])
                ops.group_sum(W["d_sum"], g[kb + "summary_attn_layer.out_proj.bias"], outer=BT, J=1, inner=1)
                ops.temporal_attention_bwd(W["sqkv"][i], W["dsmix"], W["dsqkv"], B, T, H)
                ops.gemm(W["dsqkv"], P[k + "sqkv.wT"], W["dsn"], mode=ops.EPI_BIAS_BF16)
                self._wgrad(W, W["dsqkv"], W["sum_n"][i], g[kb + "sqkv.weight"])
                ops.group_sum(W["dsqkv"], g[kb + "sqkv.bias"], outer=BT, J=1, inner=1)
                g[kb + "summary_ln.weight"].zero_(); g[kb + "summary_ln.bias"].zero_()
                ops.layernorm_bwd(W["dsn"], cp, P[k + "summary_ln.w"], W["d_cp"], rows=BT, D=D, add1=W["d_lcp"],
                                  add2=W["d_sum"], dx_bf16=W["d_cp_bf"], dw=g[kb + "summary_ln.weight"],
                                  db=g[kb + "summary_ln.bias"])
            elif e.use_local_prompts:
                # cls_proj feeds the local rows only (utils.py:176-180): d cls_proj = d(local + cls_proj rows)
                ops.layernorm(W["d_lcp"], None, None, W["d_cp_bf"], rows=BT, in_row_stride=D)      # cast
            if side is not None:
                ev_b = torch.cuda.Event()
                ev_b.record(side)
        if side is not None:
            main.wait_event(ev_b)
        # ---- cls_proj (frozen): dx[cls rows] += d_cp Wc; then refresh the bf16 copy of dx's cls rows (the LayerNorm-
        # backward kernels keep the other rows of dxb current as a by-product)
        if e.use_summary_token or e.use_local_prompts:
            ops.gemm(W["d_cp_bf"], P[k + "cls.wT"], dx_cls, mode=ops.EPI_BIAS_RESID_F32, resid=dx_cls)
        ops.layernorm(dx, None, None, dxb, rows=BT, in_row_stride=N0 * D, out_row_stride=N0 * D)

    def _bwd_tail(self, W: Dict[str, object], B: int, T: int) -> None:
        """ln_pre and the embedding sum: d time_embed[t] = sum over videos and tokens   (vision_encoder.py:108-113)."""
        e, P, g = self.enc, self.packed, self.grads
        D, N0 = e.feature_dim, int(e.num_patches)
        M = B * T * N0
        # out of place (W["gx0"]): a replayed segment must not consume its own input
        ops.layernorm_bwd(W["dx"], W["x0_pre"], P["ln_pre.w"], W["gx0"], rows=M, D=D)
        ops.group_sum(W["gx0"], g["time_embed"], outer=B, J=T, inner=N0)

    def _wgrad(self, W: Dict[str, object], dy: Tensor, x: Tensor, out: Tensor) -> None:
        """out[N, K] (f32) = dy[rows, N]^T x[rows, K] on the tensor cores: both operands transposed to [*, rows_pad]."""
        n, kk = dy.shape[1], x.shape[1]
        tA = W["tA"][:n]
        tB = W["tB"][:kk]
        ops.transpose_bf16(dy, tA)
        ops.transpose_bf16(x, tB)
        ops.gemm(tA, tB, out, mode=ops.EPI_BIAS_F32)


class TowerFunction(torch.autograd.Function):
    """autograd bridge: (video, trainable tower parameters) -> (frame_feat [BT,E], summary tokens [BT,D])."""

    @staticmethod
    def forward(ctx, trainer: TowerTrainer, video: Tensor, E: int, *params: Tensor):
        ff, st = trainer.forward_train(video)
        B, T, _, _ = _clip_dims(video)
        ctx.trainer, ctx.B, ctx.T, ctx.E = trainer, B, T, E
        ctx.generation = trainer.tape_generation
        ctx.names = [n for n, _ in trainer.trainable_named()]
        # clones: the workspace is overwritten by the next forward while autograd may still hold these outputs
        return ff[:, :E].clone(), (st.clone() if st is not None else None)      # no summary token: (frame_feat, None)

    @staticmethod
    def backward(ctx, d_ff: Optional[Tensor], d_st: Optional[Tensor]):
        tr: TowerTrainer = ctx.trainer
        grads = tr.backward(ctx.B, ctx.T, d_ff, d_st, generation=ctx.generation)
        out = []
        for (name, p), need in zip(tr.trainable_named(), ctx.needs_input_grad[3:]):
            # clones: autograd may adopt the returned tensor as .grad; the flat buffer is rewritten by the next backward
            # (gradient accumulation over micro-batches, train.py:426-485, must add, not alias)
            out.append(grads[name].view_as(p).clone() if need else None)
        return (None, None, None, *out)
