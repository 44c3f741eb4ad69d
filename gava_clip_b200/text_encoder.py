"""Text side of the logit head: CLIP text transformer + learnable / knowledge-aware prompts.

Interface and state_dict keys mirror the reference
    training/VitaCLIP_text_encoder.py:120-171 (CLIPTextEncoder), :173-379 (TextPromptLearner)
so checkpoints load unchanged.  The class text features do not depend on the video, so `TextTower`
evaluates ALL description rows of ALL classes in one batched pass and caches the result per parameter
version (the reference re-runs the 12-layer transformer once per class on every forward,
VitaCLIP_model.py:282-290).  On a CUDA device the transformer itself runs on the sm_100a kernels
(`TextTower._encode_sm100`: LayerNorm, split-bf16 3-term tcgen05 GEMMs, fp32 causal attention, fp32 QuickGELU; row (f)-1
of SURVEY.md section 8); only prompt assembly (concatenation, the tiny KAPT projector) stays in PyTorch.
"""
from __future__ import annotations

import os

from collections import OrderedDict
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .kapt_head import ContextualPromptLearner
from .tokenizer import EOT_ID, default_tokenizer, encode_with_specials, tokenize  # noqa: F401  (re-export)
from .vision_encoder import LayerNorm, QuickGELU

Tensor = torch.Tensor


class ResidualAttentionBlock(nn.Module):
    """Holder with the reference's key names (text_encoder.py:67-88): attn.in_proj_*, attn.out_proj.*, ln_1,
    mlp.c_fc, mlp.c_proj, ln_2."""

    def __init__(self, d_model: int, n_head: int, attn_mask: Optional[Tensor] = None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()), ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.n_head = n_head

    def forward(self, x: Tensor) -> Tensor:
        """x: (n, L, W) batch-first.  Causal multi-head attention + QuickGELU MLP, pre-LN."""
        n, L, W = x.shape
        h = self.ln_1(x)
        qkv = F.linear(h, self.attn.in_proj_weight, self.attn.in_proj_bias)
        q, k, v = (t.view(n, L, self.n_head, W // self.n_head).transpose(1, 2) for t in qkv.split(W, dim=-1))
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        x = x + self.attn.out_proj(o.transpose(1, 2).reshape(n, L, W))
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: Optional[Tensor] = None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: Tensor) -> Tensor:
        for blk in self.resblocks:
            x = blk(x)
        return x


class CLIPTextEncoder(nn.Module):
    def __init__(self, embed_dim: int = 512, context_length: int = 77, vocab_size: int = 49408,
                 transformer_width: int = 512, transformer_heads: int = 8, transformer_layers: int = 12):
        super().__init__()
        self.context_length = context_length
        self.vocab_size = vocab_size
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        # the reference leaves these two torch.empty (text_encoder.py:141,143) and relies on a checkpoint;
        # CLIP's own init is used here so a freshly built model is finite.
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        nn.init.normal_(self.positional_embedding, std=0.01)
        nn.init.normal_(self.text_projection, std=transformer_width ** -0.5)

    def build_attention_mask(self) -> Tensor:
        return torch.full((self.context_length, self.context_length), float("-inf")).triu_(1)

    def forward(self, prompts: Tensor, tokenized_prompts: Tensor) -> Tensor:
        """prompts (n, L, W) assembled token embeddings; returns (n, E): ln_final row at the EOT position of
        `tokenized_prompts` times text_projection (text_encoder.py:154-171)."""
        x = self.transformer(prompts + self.positional_embedding)
        x = self.ln_final(x)
        eot = torch.where(tokenized_prompts.to(x.device) == self.vocab_size - 1)[-1]
        return x[torch.arange(x.shape[0], device=x.device), eot] @ self.text_projection


class TextPromptLearner(nn.Module):
    """CoOp-style class-specific context or knowledge-aware prompts (text_encoder.py:173-379)."""

    def __init__(self, classnames: Sequence[str], text_model: CLIPTextEncoder, num_prompts: int,
                 prompts_init: str = "", CSC: bool = False, ctx_pos: str = "end", cls_type: str = "updrs",
                 knowledge_version: List[str] = ["v0"], use_descriptor: bool = False, token_wise_mlp: bool = False):
        super().__init__()
        n_cls, n_ctx = len(classnames), num_prompts
        ctx_init = prompts_init.lower()
        assert ctx_init == "" or set(ctx_init.split("_")).issubset({"split", "uni", "cntn", "disc"}), \
            "Invalid prompt initialization"
        ctx_dim = text_model.ln_final.weight.shape[0]
        self.knowledge_aware_prompt = ctx_init != ""
        self.use_descriptor = use_descriptor
        if self.knowledge_aware_prompt:
            flags = ctx_init.split("_")
            self.context_prompt_learner = ContextualPromptLearner(
                use_cntn="cntn" in flags, cntn_split="split" in flags, uni_mlp="uni" in flags, use_disc="disc" in flags,
                emb_dim=ctx_dim // 4, out_dim=ctx_dim, n_cls=n_cls, n_tokens=n_ctx, cls_type=cls_type,
                knowledge_version=knowledge_version, use_descriptor=use_descriptor, token_wise_mlp=token_wise_mlp)
            self.ctx = nn.Parameter(torch.zeros(n_cls, n_ctx, ctx_dim))
        else:
            shape = (n_cls, n_ctx, ctx_dim) if CSC else (n_ctx, ctx_dim)
            self.ctx = nn.Parameter(torch.empty(*shape))
            nn.init.normal_(self.ctx, std=0.02)
        classnames = [c.replace("_", " ") for c in classnames]
        if self.knowledge_aware_prompt:
            disc = self.context_prompt_learner.cls_disc
            texts = [[d + " " + classnames[c] for d in disc[c]] for c in range(n_cls)]
        else:
            placeholder = " ".join(["X"] * n_ctx)
            texts = [[placeholder + " " + name + "."] for name in classnames]
        self.name_lens = [len(encode_with_specials(c)) - 2 for c in classnames]
        self.tokenized_prompts = [tokenize(t) for t in texts]          # list[n_cls] of int32 (n_kv, 77)
        assert max(int(torch.where(t == EOT_ID)[-1].max()) for t in self.tokenized_prompts) <= 77
        self.n_cls, self.n_ctx = n_cls, n_ctx
        self.class_token_position = ctx_pos
        self.refresh_token_embeddings(text_model)

    def refresh_token_embeddings(self, text_model: CLIPTextEncoder) -> None:
        """Snapshot the frozen SOS / class-text token embeddings (plain lists, not buffers, as the reference:
        text_encoder.py:283-306).  Taken at construction, i.e. after `backbone_path` was loaded; call again only
        if `textual.token_embedding` is replaced afterwards."""
        self.token_prefix, self.token_suffix = [], []
        with torch.no_grad():
            for tok in self.tokenized_prompts:
                emb = text_model.token_embedding(tok.to(text_model.token_embedding.weight.device))
                self.token_prefix.append(emb[:, :1, :])
                # KAPT drops the LAST n_ctx positions (text shifts right by n_ctx, EOT index is NOT shifted:
                # SURVEY.md 8a quirk 1); placeholder mode drops the "X" positions.
                self.token_suffix.append(emb[:, 1:-self.n_ctx, :] if self.knowledge_aware_prompt
                                         else emb[:, 1 + self.n_ctx:, :])
        self._embed_version = getattr(self, "_embed_version", 0) + 1

    def forward(self) -> List[Tensor]:
        if self.class_token_position != "end":
            raise NotImplementedError(f"Unsupported class token position: {self.class_token_position}")
        if self.knowledge_aware_prompt:
            ctx = self.context_prompt_learner(self.ctx)               # list[n_cls] of (n_kv, n_ctx, W)
        else:
            if self.ctx.dim() != 3:
                raise NotImplementedError("generic (non class-specific) context is broken in the reference "
                                          "(text_encoder.py:315: 2-D ctx cannot be concatenated); use text_prompt_CSC=True")
            ctx = [self.ctx[c].unsqueeze(0) for c in range(self.n_cls)]
        dev = self.ctx.device
        out = []
        for c in range(self.n_cls):
            pre, suf = self.token_prefix[c].to(dev), self.token_suffix[c].to(dev)
            out.append(torch.cat([pre, ctx[c].expand(pre.shape[0], -1, -1), suf], dim=-2))
        return out


class TextTower:
    """Batched, cached evaluation of all class description rows."""

    def __init__(self, textual: CLIPTextEncoder, prompt_learner: TextPromptLearner):
        self.textual, self.prompt_learner = textual, prompt_learner
        self._sig = None
        self._cache: Optional[Tuple[Tensor, Tensor, Tensor, List[int]]] = None
        self.native_train = os.environ.get("GAVA_TEXT_NATIVE_TRAIN", "1") != "0"   # text_train.TextTrainer in training

    def _signature(self, device):
        ps = list(self.textual.parameters()) + list(self.prompt_learner.parameters())
        return (str(device), self.prompt_learner._embed_version) + tuple((p.data_ptr(), p._version) for p in ps)

    # ------------------------------------------------------------------ sm_100a path
    def _pack(self, device) -> dict:
        """hi/lo bf16 pairs of the text transformer's GEMM weights (fp32-accuracy products: the features are computed
        once per parameter version, so the 3x GEMM cost is irrelevant and the logits keep their full error budget)."""
        key = tuple((p.data_ptr(), p._version) for p in self.textual.parameters())
        if getattr(self, "_pack_key", None) == (key, str(device)):
            return self._packed
        t = self.textual

        def pair(w):
            w = w.detach().to(device, torch.float32).contiguous()
            hi = w.to(torch.bfloat16)
            return hi, (w - hi.float()).to(torch.bfloat16)

        f32 = lambda v: v.detach().to(device, torch.float32).contiguous()
        P = {"blocks": [], "pos": f32(t.positional_embedding), "lnf": (f32(t.ln_final.weight), f32(t.ln_final.bias)),
             "proj": pair(t.text_projection.t())}
        for blk in t.transformer.resblocks:
            P["blocks"].append({
                "ln1": (f32(blk.ln_1.weight), f32(blk.ln_1.bias)), "ln2": (f32(blk.ln_2.weight), f32(blk.ln_2.bias)),
                "in": pair(blk.attn.in_proj_weight), "in_b": f32(blk.attn.in_proj_bias),
                "out": pair(blk.attn.out_proj.weight), "out_b": f32(blk.attn.out_proj.bias),
                "fc": pair(blk.mlp.c_fc.weight), "fc_b": f32(blk.mlp.c_fc.bias),
                "pj": pair(blk.mlp.c_proj.weight), "pj_b": f32(blk.mlp.c_proj.bias)})
        self._packed, self._pack_key = P, (key, str(device))
        return P

    def _encode_sm100(self, prompts: Tensor, tokenized: Tensor) -> Tensor:
        """CLIPTextEncoder.forward (text_encoder.py:154-171) for all R description rows at once on the sm_100a kernels:
        LayerNorm, split-bf16 3-term tcgen05 GEMMs, fp32 causal attention, fp32 QuickGELU."""
        from . import ops
        t = self.textual
        R, L, Wd = prompts.shape
        H = t.transformer.resblocks[0].n_head
        if Wd // H != 64 or Wd % 128 != 0:
            raise NotImplementedError("text tower kernels are specialised for head_dim 64 and width % 128 == 0")
        dev = prompts.device
        P = self._pack(dev)
        M = R * L
        x = (prompts + P["pos"]).reshape(M, Wd).contiguous()
        z = lambda *s, dt=torch.float32: torch.empty(*s, device=dev, dtype=dt)
        tn, qkv, att, h = z(M, Wd), z(M, 3 * Wd), z(M, Wd), z(M, 4 * Wd)
        hi, lo = z(M * 4 * Wd, dt=torch.bfloat16), z(M * 4 * Wd, dt=torch.bfloat16)

        def mm(a32, rows, K, w, out, bias=None, resid=None):
            ah, al = hi[: rows * K].view(rows, K), lo[: rows * K].view(rows, K)
            ops.split_bf16(a32, ah, al, rows=rows, D=K)
            ops.gemm3(ah, al, w[0], w[1], out, bias=bias, resid=resid)

        for b in P["blocks"]:
            ops.layernorm(x, b["ln1"][0], b["ln1"][1], tn, rows=M, in_row_stride=Wd)
            mm(tn, M, Wd, b["in"], qkv, bias=b["in_b"])
            ops.causal_attention_f32(qkv, att, n_seq=R, L=L, H=H)
            mm(att, M, Wd, b["out"], x, bias=b["out_b"], resid=x)
            ops.layernorm(x, b["ln2"][0], b["ln2"][1], tn, rows=M, in_row_stride=Wd)
            mm(tn, M, Wd, b["fc"], h, bias=b["fc_b"])
            ops.quickgelu_f32(h)
            mm(h, M, 4 * Wd, b["pj"], x, bias=b["pj_b"], resid=x)
        eot = torch.where(tokenized == t.vocab_size - 1)[-1]                     # :169 (un-shifted ids: KAPT quirk kept)
        rows = x.view(R, L, Wd)[torch.arange(R, device=dev), eot].contiguous()    # gather: memory plumbing
        rn = z(R, Wd)
        ops.layernorm(rows, P["lnf"][0], P["lnf"][1], rn, rows=R, in_row_stride=Wd)
        E = t.text_projection.shape[1]
        ep = (E + 127) // 128 * 128
        w_hi, w_lo = P["proj"]
        if ep != E:
            w_hi = torch.cat([w_hi, w_hi.new_zeros(ep - E, Wd)]); w_lo = torch.cat([w_lo, w_lo.new_zeros(ep - E, Wd)])
        out = z(R, ep)
        mm(rn, R, Wd, (w_hi, w_lo), out)
        return out[:, :E].contiguous()

    def _tokens_on(self, device) -> Tensor:
        """concatenated token ids on `device`, cached: the SAME tensor every step, so consumers can key per-tokenisation
        state (EOT row indices, which need a host sync to compute) on it"""
        toks = self.prompt_learner.tokenized_prompts
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in toks)
        if getattr(self, "_tok_cache", (None, None))[0] != key:
            self._tok_cache = (key, torch.cat([t.to(device) for t in toks], dim=0))
        return self._tok_cache[1]

    def features(self, device: torch.device, need_grad: bool = False):
        """-> (t_hat (R,E) L2-normalised rows, class_offsets int32 (n_cls+1), text_features (n_cls,E), counts)."""
        sig = self._signature(device)
        if not need_grad and self._cache is not None and sig == self._sig:
            return self._cache
        with torch.set_grad_enabled(need_grad):
            prompts = self.prompt_learner()
            counts = [p.shape[0] for p in prompts]
            tok = self._tokens_on(device)
            if device.type == "cuda" and not need_grad:
                raw = self._encode_sm100(torch.cat(prompts, dim=0).to(device).float(), tok)
            elif device.type == "cuda" and self.native_train and not any(p.requires_grad for p in self.textual.parameters()):
                # fine-tune step: frozen transformer forward + input-gradient backward on the sm_100a kernels
                from .text_train import TextTrainer, TextTrainFunction
                if getattr(self, "_trainer", None) is None or self._trainer.device != device:
                    self._trainer = TextTrainer(self.textual, device)
                raw = TextTrainFunction.apply(self._trainer, torch.cat(prompts, dim=0).to(device), tok)
            elif device.type == "cuda":
                # No PyTorch fallback on the GPU: the sm_100a text backward covers what the reference trains (prompt
                # learner / KAPT projector through a FROZEN text transformer, VitaCLIP_model.py:230-239).
                bad = [n for n, p in self.textual.named_parameters() if p.requires_grad][:4]
                raise NotImplementedError(
                    "gava_clip_b200 has no PyTorch fallback on CUDA: the text side of the fine-tune step runs on the sm_100a "
                    "kernels with the text transformer frozen (the reference's recipe); "
                    + (f"these textual parameters require grad: {bad}" if bad else
                       "GAVA_TEXT_NATIVE_TRAIN=0 disabled that path"))
            else:       # host-side (CPU) tests of the module logic only: the module's own PyTorch forward
                raw = self.textual(torch.cat(prompts, dim=0).to(device), tok).float()
            t_hat = raw / raw.norm(dim=-1, keepdim=True)
            offs = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32, device=device)
            means = torch.stack([c.mean(0) for c in t_hat.split(counts, dim=0)])
            text_features = means / means.norm(dim=-1, keepdim=True)          # VitaCLIP_model.py:290-293
        out = (t_hat.contiguous(), offs, text_features, counts)
        if not need_grad:
            self._cache, self._sig = out, sig
        return out
