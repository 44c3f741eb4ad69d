"""Text side of the fine-tune step on the sm_100a kernels ("text encoder fwd+bwd every step").

The CLIP text transformer is frozen (VitaCLIP_model.py:237-239); what trains is the prompt learner in front of it (class
context / KAPT projector), so every step needs the transformer's forward for the R description rows and the gradient
with respect to its INPUT embeddings.  `TextTrainer` runs both on the same kernels as the video tower:

  forward   x = prompts + positional_embedding; per block: LayerNorm -> in_proj GEMM (f32 q|k|v) -> causal attention (f32)
            -> out_proj GEMM (+residual) -> LayerNorm -> c_fc GEMM (bf16 pre-activations kept) -> QuickGELU -> c_proj GEMM
            (+residual); ln_final on the EOT rows -> text_projection GEMM          (VitaCLIP_text_encoder.py:67-88,154-171)
  backward  the same chain reversed: dgrad GEMMs on W^T, DGELU epilogue, LayerNorm backward, causal attention backward.

The FORWARD uses the fp32-accuracy products of the inference text path (split-bf16 3-term tcgen05 GEMMs, fp32
activations): the text features feed every logit, and at R x 77 rows the 3x GEMM cost is microseconds, so the logits keep
the error budget of the 2e-2 parity bound for the video tower.  The BACKWARD runs plain bf16 operands with fp32 accumulation
(gradient noise of 2^-8 is far inside what the reference's fp16 autocast, train.py:441, produces).  `TextTrainFunction`
exposes the pair to autograd: prompt assembly (a few concatenations and the tiny KAPT projector MLP) stays in PyTorch in
front of it, the cosine head behind it.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from . import ops

Tensor = torch.Tensor


class TextTrainer:
    def __init__(self, textual, device: torch.device):
        self.textual, self.device = textual, device
        self._pack_key = None
        self.P: Dict[str, object] = {}
        self.ws: Dict[tuple, Dict[str, object]] = {}
        self.graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self.use_graph = os.environ.get("GAVA_CUDA_GRAPH", "1") != "0"

    # ------------------------------------------------------------------ frozen weights: bf16 [N,K] and transposed copies
    def pack(self) -> None:
        t, dev = self.textual, self.device
        key = tuple((p.data_ptr(), p._version) for p in t.parameters())
        if key == self._pack_key:
            return
        bf = lambda w: w.detach().to(dev, torch.bfloat16).contiguous()
        bft = lambda w: w.detach().t().to(dev, torch.bfloat16).contiguous()

        def pair(w):      # [w_hi | w_hi | w_lo] along K: operand of the K-concatenated 3-term forward products
            return ops.pack3(w.detach().to(dev, torch.float32))

        f32 = lambda v: v.detach().to(dev, torch.float32).contiguous()
        Wd, E = t.text_projection.shape
        Ep = (E + 127) // 128 * 128
        pT32 = torch.zeros(Ep, Wd, device=dev, dtype=torch.float32)      # forward operand: text_projection^T, rows padded
        pT32[:E] = t.text_projection.detach().t().to(dev, torch.float32)
        pT = pair(pT32)
        pw = torch.zeros(Wd, Ep, device=dev, dtype=torch.bfloat16)       # backward operand: text_projection, K padded
        pw[:, :E] = bf(t.text_projection)
        P: Dict[str, object] = {"pos": f32(t.positional_embedding), "lnf": (f32(t.ln_final.weight), f32(t.ln_final.bias)),
                                "proj.wT": pT, "proj.w": pw, "blocks": []}
        for blk in t.transformer.resblocks:
            P["blocks"].append({
                "ln1": (f32(blk.ln_1.weight), f32(blk.ln_1.bias)), "ln2": (f32(blk.ln_2.weight), f32(blk.ln_2.bias)),
                "in.w": pair(blk.attn.in_proj_weight), "in.wT": bft(blk.attn.in_proj_weight), "in.b": f32(blk.attn.in_proj_bias),
                "out.w": pair(blk.attn.out_proj.weight), "out.wT": bft(blk.attn.out_proj.weight), "out.b": f32(blk.attn.out_proj.bias),
                "fc.w": pair(blk.mlp.c_fc.weight), "fc.wT": bft(blk.mlp.c_fc.weight), "fc.b": f32(blk.mlp.c_fc.bias),
                "pj.w": pair(blk.mlp.c_proj.weight), "pj.wT": bft(blk.mlp.c_proj.weight), "pj.b": f32(blk.mlp.c_proj.bias)})
        self.P, self._pack_key = P, key
        self.graphs.clear()      # captured graphs hold pointers into the old packed tensors
        self.ws.clear()

    def workspace(self, R: int, L: int) -> Dict[str, object]:
        key = (R, L)
        if key in self.ws:
            return self.ws[key]
        t, dev = self.textual, self.device
        Wd = t.text_projection.shape[0]
        Ep = self.P["proj.wT"].shape[0]
        nl = len(t.transformer.resblocks)
        M = R * L
        bf, f32 = torch.bfloat16, torch.float32
        z = lambda *s, dt: torch.empty(*s, device=dev, dtype=dt)
        W = {"x_in": [z(M, Wd, dt=f32) for _ in range(nl + 1)], "x_mid": [z(M, Wd, dt=f32) for _ in range(nl)],
             "qkv": [z(M, 3 * Wd, dt=f32) for _ in range(nl)], "a1": [z(M, 4 * Wd, dt=bf) for _ in range(nl)],
             "tn": z(M, Wd, dt=f32), "att": z(M, Wd, dt=f32), "h": z(M, 4 * Wd, dt=f32),
             "a3": z(M * 12 * Wd, dt=bf),
             "rows": z(R, Wd, dt=f32), "rn": z(R, Wd, dt=f32), "feat": z(R, Ep, dt=f32),
             "dx": z(M, Wd, dt=f32), "dxb": z(M, Wd, dt=bf), "g4": z(M, 4 * Wd, dt=bf), "gD": z(M, Wd, dt=bf),
             "dqkv": z(M, 3 * Wd, dt=bf), "dfeat": z(R, Ep, dt=f32), "dfeat_bf": z(R, Ep, dt=bf), "drn": z(R, Wd, dt=bf),
             "drows": z(R, Wd, dt=f32),
             "prompts_in": z(M, Wd, dt=f32), "pos_rep": self.P["pos"][:L].repeat(R, 1).contiguous(),
             "row_idx": torch.zeros(R, device=dev, dtype=torch.long)}
        W["dfeat"].zero_()
        self.ws[key] = W
        return W

    # ------------------------------------------------------------------ forward with tape
    def _forward_body(self, R: int, L: int) -> None:
        """W["prompts_in"] (f32 [R*L, W]) -> W["feat"]; every buffer is static, so the launch sequence can be replayed."""
        t = self.textual
        P, W = self.P, self.workspace(R, L)
        Wd = t.text_projection.shape[0]
        H = t.transformer.resblocks[0].n_head
        M = R * L
        torch.add(W["prompts_in"], W["pos_rep"], out=W["x_in"][0])                          # :157 (memory plumbing)

        def mm(a32: Tensor, rows: int, K: int, w: Tensor, out: Tensor, bias=None, resid=None) -> None:
            """out(f32) = [resid +] a32 @ (w_hi + w_lo)^T + bias as a 3-term split-bf16 product in ONE GEMM (K' = 3K)."""
            a3 = W["a3"][: rows * 3 * K].view(rows, 3 * K)
            ops.split3_bf16(a32, a3, rows=rows, D=K)
            if resid is not None:
                ops.gemm(a3, w, out, mode=ops.EPI_BIAS_RESID_F32, bias=bias, resid=resid)
            else:
                ops.gemm(a3, w, out, mode=ops.EPI_BIAS_F32, bias=bias)

        for i, b in enumerate(P["blocks"]):
            xin, xmid, xnext = W["x_in"][i], W["x_mid"][i], W["x_in"][i + 1]
            ops.layernorm(xin, b["ln1"][0], b["ln1"][1], W["tn"], rows=M, in_row_stride=Wd)
            mm(W["tn"], M, Wd, b["in.w"], W["qkv"][i], bias=b["in.b"])
            ops.causal_attention_f32(W["qkv"][i], W["att"], n_seq=R, L=L, H=H)
            mm(W["att"], M, Wd, b["out.w"], xmid, bias=b["out.b"], resid=xin)
            ops.layernorm(xmid, b["ln2"][0], b["ln2"][1], W["tn"], rows=M, in_row_stride=Wd)
            mm(W["tn"], M, Wd, b["fc.w"], W["h"], bias=b["fc.b"])
            ops.layernorm(W["h"].view(4 * M, Wd), None, None, W["a1"][i].view(4 * M, Wd), rows=4 * M, in_row_stride=Wd)  # keep bf16 pre-activations
            ops.quickgelu_f32(W["h"])
            mm(W["h"], M, 4 * Wd, b["pj.w"], xnext, bias=b["pj.b"], resid=xmid)
        torch.index_select(W["x_in"][len(P["blocks"])], 0, W["row_idx"], out=W["rows"])     # EOT gather: memory plumbing
        ops.layernorm(W["rows"], P["lnf"][0], P["lnf"][1], W["rn"], rows=R, in_row_stride=Wd)
        mm(W["rn"], R, Wd, P["proj.wT"], W["feat"])

    def _replay(self, kind: str, R: int, L: int, body) -> None:
        """Run `body` eagerly the first time for a shape (kernel attributes, allocator), then from a captured CUDA graph:
        the text side is ~480 latency-bound launches per step, replayed back-to-back instead of issued one by one."""
        if not self.use_graph or torch.cuda.is_current_stream_capturing():
            body(R, L)
            return
        key = (kind, R, L)
        g = self.graphs.get(key)
        if g is None:
            body(R, L)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):   # a DataLoader thread may be busy with CUDA calls
                body(R, L)
            self.graphs[key] = g
        g.replay()

    def forward_train(self, prompts: Tensor, tokenized: Tensor) -> Tensor:
        """prompts f32 [R, L, W] (assembled token embeddings), tokenized int [R, L] -> raw text features f32 [R, E]."""
        t = self.textual
        R, L, Wd = prompts.shape
        H = t.transformer.resblocks[0].n_head
        if Wd // H != 64 or Wd % 128 != 0 or L > 128:
            raise NotImplementedError("text tower kernels are specialised for head_dim 64, width % 128 == 0, <= 128 tokens")
        self.pack()
        W = self.workspace(R, L)
        tok_key = (tokenized.data_ptr(), tokenized._version, R, L)
        if W.get("tok_key") != tok_key:
            eot = torch.where(tokenized == t.vocab_size - 1)[-1]                  # :169 (un-shifted ids: KAPT quirk kept)
            W["row_idx"].copy_(torch.arange(R, device=prompts.device) * L + eot)
            W["tok_key"] = tok_key
        W["prompts_in"].copy_(prompts.reshape(R * L, Wd))
        self._replay("fwd", R, L, self._forward_body)
        return W["feat"][:, : t.text_projection.shape[1]]

    # ------------------------------------------------------------------ backward
    def _backward_body(self, R: int, L: int) -> None:
        """W["dfeat"] (f32 [R, Ep], zero-padded) -> W["dx"] = d(prompts)."""
        t = self.textual
        P, W = self.P, self.workspace(R, L)
        Wd = t.text_projection.shape[0]
        H = t.transformer.resblocks[0].n_head
        M = R * L
        dx, dxb, gD = W["dx"], W["dxb"], W["gD"]
        Ep = W["dfeat"].shape[1]
        ops.layernorm(W["dfeat"], None, None, W["dfeat_bf"], rows=R, in_row_stride=Ep)
        ops.gemm(W["dfeat_bf"], P["proj.w"], W["drn"], mode=ops.EPI_BIAS_BF16)
        ops.layernorm_bwd(W["drn"], W["rows"], P["lnf"][0], W["drows"], rows=R, D=Wd)
        dx.zero_()
        dx.index_copy_(0, W["row_idx"], W["drows"])                               # scatter to the EOT rows
        ops.layernorm(dx, None, None, dxb, rows=M, in_row_stride=Wd)              # bf16 copy; kept current by the LN backward
        for i in reversed(range(len(P["blocks"]))):
            b = P["blocks"][i]
            ops.gemm(dxb, b["pj.wT"], W["g4"], mode=ops.EPI_DGELU_BF16, aux=W["a1"][i])
            ops.gemm(W["g4"], b["fc.wT"], gD, mode=ops.EPI_BIAS_BF16)
            ops.layernorm_bwd(gD, W["x_mid"][i], b["ln2"][0], dx, rows=M, D=Wd, add1=dx, dx_bf16=dxb)
            ops.gemm(dxb, b["out.wT"], gD, mode=ops.EPI_BIAS_BF16)
            ops.causal_attention_bwd(W["qkv"][i], gD, W["dqkv"], n_seq=R, L=L, H=H)
            ops.gemm(W["dqkv"], b["in.wT"], gD, mode=ops.EPI_BIAS_BF16)
            ops.layernorm_bwd(gD, W["x_in"][i], b["ln1"][0], dx, rows=M, D=Wd, add1=dx, dx_bf16=dxb)

    def backward(self, R: int, L: int, d_feat: Tensor) -> Tensor:
        """d(raw features) [R, E] -> d(prompts) f32 [R, L, W] (a workspace view; valid until the next backward)."""
        W = self.workspace(R, L)
        W["dfeat"][:, : d_feat.shape[1]].copy_(d_feat)                            # padding columns stay zero
        self._replay("bwd", R, L, self._backward_body)
        return W["dx"].view(R, L, -1)


class TextTrainFunction(torch.autograd.Function):
    """autograd bridge: assembled prompt embeddings [R, L, W] -> raw text features [R, E] through the frozen transformer."""

    @staticmethod
    def forward(ctx, trainer: TextTrainer, prompts: Tensor, tokenized: Tensor):
        feat = trainer.forward_train(prompts.contiguous().float(), tokenized)
        ctx.trainer, ctx.shape = trainer, tuple(prompts.shape)
        return feat.clone()

    @staticmethod
    def backward(ctx, d_feat: Tensor):
        R, L, _ = ctx.shape
        return None, ctx.trainer.backward(R, L, d_feat.contiguous().float()).clone(), None
