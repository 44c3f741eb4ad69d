"""CPU oracle for the Vita-CLIP / GaVA-CLIP hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32 restatement of the reference algorithm for the
path named in BASELINE.json (video encoder forward + contrastive logit head).
It is *never* imported by the product package (`gava_clip_b200`); only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may use it, and only as the checker / CPU baseline.

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against the reference *itself*, executed in the build container
by `oracle/make_golden.py` (imports `/root/reference`, same weights, same clips);
`tests/test_oracle_golden.py` re-checks the oracle against those committed
outputs (`tests/golden/*.pt`) wherever the repo runs.

Every function cites the reference lines it restates (paths relative to the
reference root, `training/...`).  State-dict keys are the reference's own.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------- #
# primitives
# --------------------------------------------------------------------------- #
def quick_gelu(x: Tensor) -> Tensor:
    """x * sigmoid(1.702 x)   -- VitaCLIP_vision_encoder_utils.py:18-20"""
    return x * torch.sigmoid(1.702 * x)


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """fp32 LayerNorm, cast back to input dtype -- vision_encoder_utils.py:22-28"""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps).to(x.dtype)


def linear(x: Tensor, sd: SD, name: str) -> Tensor:
    w = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    return F.linear(x, w, b)


def mhsa(sd: SD, prefix: str, q_in: Tensor, k_in: Tensor, v_in: Tensor, num_heads: int) -> Tensor:
    """Generalised attention -- vision_encoder_utils.py:61-81.

    separate q/k/v projections, q scaled by dh^-0.5 *before* the dot product,
    softmax over the key axis, out projection.
    """
    n, lq, _ = q_in.shape
    lk = k_in.shape[1]
    q = linear(q_in, sd, prefix + "q_proj")
    k = linear(k_in, sd, prefix + "k_proj")
    v = linear(v_in, sd, prefix + "v_proj")
    dh = q.shape[-1] // num_heads
    q = q.view(n, lq, num_heads, dh)
    k = k.view(n, lk, num_heads, dh)
    v = v.view(n, lk, num_heads, v.shape[-1] // num_heads)
    aff = torch.einsum("nqhc,nkhc->nqkh", q / (dh ** 0.5), k)
    aff = aff.softmax(dim=-2)
    mix = torch.einsum("nqlh,nlhc->nqhc", aff, v)
    return linear(mix.flatten(-2), sd, prefix + "out_proj")


# --------------------------------------------------------------------------- #
# vision tower
# --------------------------------------------------------------------------- #
def patch_embed(sd: SD, prefix: str, frames: Tensor) -> Tensor:
    """Conv2d(k=s=patch)+bias -> (BT, Np, D) -- vision_encoder_utils.py:215-219"""
    w = sd[prefix + "patch_embed.proj.weight"]
    b = sd[prefix + "patch_embed.proj.bias"]
    p = w.shape[-1]
    y = F.conv2d(frames, w, b, stride=p)
    return y.flatten(2).transpose(1, 2)


def time_embed_for(sd: SD, prefix: str, T: int) -> Tensor:
    """nearest-neighbour resize of time_embed when T differs -- vision_encoder.py:91-97"""
    te = sd[prefix + "time_embed"]
    if T == te.shape[0]:
        return te
    src = te.unsqueeze(0).transpose(1, 2)
    return F.interpolate(src, size=(T,), mode="nearest").transpose(1, 2).squeeze(0)


def encoder_layer(
    sd: SD, prefix: str, x: Tensor, num_frames: int, num_heads: int,
    use_summary: bool = True, use_local: bool = True,
) -> Tuple[Tensor, Optional[Tensor]]:
    """One TransformerEncoderLayer -- vision_encoder_utils.py:155-203.

    x: (BT, 1+G+Np, D) with token order [cls | global | patches].
    """
    BT, _, D = x.shape
    T = num_frames
    B = BT // T
    summary = None
    n_local = 0
    if use_summary or use_local:
        cls_proj = linear(x[:, 0, :].reshape(B, T, D), sd, prefix + "cls_proj")       # :160-165
    if use_summary:
        s_norm = layer_norm(cls_proj, sd[prefix + "summary_ln.weight"], sd[prefix + "summary_ln.bias"])
        summary = cls_proj + mhsa(sd, prefix + "summary_attn_layer.", s_norm, s_norm, s_norm, num_heads)  # :168-170
        x = torch.cat([x, summary.reshape(BT, 1, D)], dim=1)                              # :171-172
    if use_local:
        lp = sd[prefix + "local_prompts"].expand(B, -1, -1) + cls_proj                    # :176-185
        lp = lp.repeat_interleave(T, dim=0)                                               # :187
        n_local = lp.shape[1]
        x = torch.cat([x[:, :1], lp, x[:, 1:]], dim=1)                                    # :188
    xn = layer_norm(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"])
    x = x + mhsa(sd, prefix + "attn.", xn, xn, xn, num_heads)                             # :190-191
    if use_summary:
        x = x[:, :-1]                                                                     # :194-195
    if use_local:
        x = torch.cat([x[:, :1], x[:, 1 + n_local:]], dim=1)                              # :196-197
    h = layer_norm(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"])
    h = linear(quick_gelu(linear(h, sd, prefix + "mlp.fc1")), sd, prefix + "mlp.fc2")     # :110-115
    x = x + h                                                                             # :199
    return x, summary


def vision_encoder(
    sd: SD, video: Tensor, *, num_frames: int, num_heads: int, num_layers: int,
    num_global_prompts: int = 8, prefix: str = "visual.", taps: Optional[dict] = None,
    use_summary: bool = True, use_local: bool = True,
) -> Tuple[Tensor, Optional[Tensor]]:
    """CLIPVisionEncoder.forward -- VitaCLIP_vision_encoder.py:102-132 (use_global_prompts=True: its other branch crashes).

    video: (B, 3, T, H, W) fp32.  Returns (cls_x (B,E), summary (B,D)).
    `taps`, if given, receives per-layer cls rows and summary tokens (debug/parity aids).
    """
    B, C, T, H, W = video.shape
    assert T == num_frames, "the reference blocks hard-code num_frames (utils.py:161)"
    frames = video.permute(0, 2, 1, 3, 4).flatten(0, 1)                                   # :105
    x = patch_embed(sd, prefix, frames)                                                   # :107
    cls = sd[prefix + "cls_token"].view(1, 1, -1).expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1)                                                        # :108
    x = x + sd[prefix + "pos_embed"]                                                      # :110
    te = time_embed_for(sd, prefix, T)                                                    # :111 (+ :86-100)
    x = (x.view(B, T, x.shape[1], -1) + te.view(1, T, 1, -1)).flatten(0, 1)
    x = layer_norm(x, sd[prefix + "ln_pre.weight"], sd[prefix + "ln_pre.bias"])           # :113
    if taps is not None:
        taps["embed_cls"] = x[:, 0].clone()
        taps["embed"] = x.clone()
    G = num_global_prompts
    summary = None
    for i in range(num_layers):                                                           # :116-121
        gp = sd[prefix + "global_prompts"][i].expand(B * T, -1, -1)
        x = torch.cat([x[:, :1], gp, x[:, 1:]], dim=1)
        x, summary = encoder_layer(sd, f"{prefix}blocks.{i}.", x, num_frames, num_heads, use_summary, use_local)
        x = torch.cat([x[:, :1], x[:, G + 1:]], dim=1)
        if taps is not None:
            taps.setdefault("layer_cls", []).append(x[:, 0].clone())
            if summary is not None:
                taps.setdefault("layer_summary", []).append(summary.clone())
            taps.setdefault("layer_patch_mean", []).append(x[:, 1:].mean(dim=(1,)).clone())
    cls_x = layer_norm(x[:, 0], sd[prefix + "ln_post.weight"], sd[prefix + "ln_post.bias"])  # :126
    cls_x = cls_x @ sd[prefix + "proj"]                                                   # :127
    cls_x = cls_x.view(B, T, -1).mean(dim=1)                                              # :128
    if summary is not None:
        summary = summary.mean(dim=1)                                                     # :129-130
    return cls_x, summary


# --------------------------------------------------------------------------- #
# text side (producer of the head operand; video independent)
# --------------------------------------------------------------------------- #
def text_resblock(sd: SD, prefix: str, x: Tensor, heads: int, mask: Tensor) -> Tensor:
    """ResidualAttentionBlock with nn.MultiheadAttention -- VitaCLIP_text_encoder.py:67-88.

    x: (N, L, W).  MultiheadAttention semantics: packed in_proj, q scaled by dh^-0.5,
    additive causal mask, out_proj.
    """
    n, l, w = x.shape
    dh = w // heads
    h = layer_norm(x, sd[prefix + "ln_1.weight"], sd[prefix + "ln_1.bias"])
    qkv = F.linear(h, sd[prefix + "attn.in_proj_weight"], sd[prefix + "attn.in_proj_bias"])
    q, k, v = qkv.split(w, dim=-1)
    q = q.view(n, l, heads, dh).transpose(1, 2)
    k = k.view(n, l, heads, dh).transpose(1, 2)
    v = v.view(n, l, heads, dh).transpose(1, 2)
    att = (q * dh ** -0.5) @ k.transpose(-1, -2) + mask
    att = att.softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(n, l, w)
    x = x + F.linear(o, sd[prefix + "attn.out_proj.weight"], sd[prefix + "attn.out_proj.bias"])
    h = layer_norm(x, sd[prefix + "ln_2.weight"], sd[prefix + "ln_2.bias"])
    h = F.linear(h, sd[prefix + "mlp.c_fc.weight"], sd[prefix + "mlp.c_fc.bias"])
    h = F.linear(quick_gelu(h), sd[prefix + "mlp.c_proj.weight"], sd[prefix + "mlp.c_proj.bias"])
    return x + h


def text_encoder(
    sd: SD, prompts: Tensor, tokenized: Tensor, *, heads: int, layers: int,
    vocab_size: int = 49408, prefix: str = "textual.",
) -> Tensor:
    """CLIPTextEncoder.forward -- VitaCLIP_text_encoder.py:154-171.

    prompts: (n, 77, W) token embeddings (already prompt-assembled);
    tokenized: (n, 77) ids, used only to locate the EOT position (:169).
    """
    n, l, w = prompts.shape
    x = prompts + sd[prefix + "positional_embedding"]
    mask = torch.full((l, l), float("-inf"), dtype=x.dtype, device=x.device).triu_(1)     # :146-152
    for i in range(layers):
        x = text_resblock(sd, f"{prefix}transformer.resblocks.{i}.", x, heads, mask)
    x = layer_norm(x, sd[prefix + "ln_final.weight"], sd[prefix + "ln_final.bias"])
    eot = torch.where(tokenized == vocab_size - 1)[-1]                                    # :169
    return x[torch.arange(n, device=x.device), eot] @ sd[prefix + "text_projection"]


def kapt_context(sd: SD, ctx: Tensor, cntn_embeds: Sequence[Tensor], prefix: str) -> List[Tensor]:
    """ContextualPromptLearner.forward, class-wise uni-MLP branch -- kapt_head.py:177-199.

    ctx: (n_cls, n_ctx, W); cntn_embeds[c]: (n_kv, 768).  Returns per-class (n_kv, n_ctx, W).
    """
    out = []
    n_ctx = ctx.shape[1]
    for c, emb in enumerate(cntn_embeds):
        w0 = sd[f"{prefix}projector.{c}.0.weight"]
        w2 = sd[f"{prefix}projector.{c}.2.weight"]
        e = F.linear(F.relu(F.linear(emb, w0)), w2).unsqueeze(1).expand(-1, n_ctx, -1)
        out.append(ctx[c].unsqueeze(0) + e)
    return out


def assemble_prompts(
    sd: SD, tokenized: Sequence[Tensor], n_ctx: int, *, kapt: bool,
    cntn_embeds: Optional[Sequence[Tensor]] = None,
) -> List[Tensor]:
    """TextPromptLearner.__init__/forward -- text_encoder.py:283-300, 310-332.

    prefix = embedding of token 0 (SOS); suffix = embedding of tokens [1+n_ctx:] in the
    placeholder ("X X ... name.") mode, or tokens [1:77-n_ctx] in KAPT mode (:296-298);
    prompt = [prefix | ctx | suffix].
    """
    emb_w = sd["textual.token_embedding.weight"]
    ctx = sd["prompt_learner.ctx"]
    if kapt:
        ctx_per_cls = kapt_context(sd, ctx, cntn_embeds, "prompt_learner.context_prompt_learner.")
    prompts = []
    for c, tok in enumerate(tokenized):
        emb = emb_w[tok.long()]                                   # (n_kv, 77, W)
        pre = emb[:, :1]
        if kapt:
            suf = emb[:, 1:-n_ctx]
            cc = ctx_per_cls[c]
        else:
            suf = emb[:, 1 + n_ctx:]
            cc = (ctx[c] if ctx.dim() == 3 else ctx).unsqueeze(0).expand(emb.shape[0], -1, -1)
        prompts.append(torch.cat([pre, cc, suf], dim=1))
    return prompts


# --------------------------------------------------------------------------- #
# logit head
# --------------------------------------------------------------------------- #
def logit_head(
    video_features: Tensor, class_text_features: Sequence[Tensor], logit_scale: Tensor,
    logit_bias: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor]:
    """Prompt-learning branch of VitaCLIP.forward -- VitaCLIP_model.py:248-293, 308-309.

    class_text_features[c]: (n_kv, E) un-normalised text features of class c.
    logits[:, c] = mean_k( exp(logit_scale) * v_hat . t_hat_{c,k} );  also returns the
    side-effect tensor `text_features` (normalised per-class mean, :290-293).
    """
    scale = logit_scale.exp()
    v = video_features / video_features.norm(dim=-1, keepdim=True)
    cols, means = [], []
    for t in class_text_features:
        t = t / t.norm(dim=-1, keepdim=True)
        cols.append((scale * v @ t.t()).mean(-1, keepdim=True))
        means.append(t.mean(0, keepdim=True))
    logits = torch.cat(cols, dim=-1)
    tf = torch.cat(means, dim=0)
    tf = tf / tf.norm(dim=-1, keepdim=True)
    if logit_bias is not None:
        logits = logits + logit_bias
    return logits, tf


def zeroshot_head(video_features: Tensor, text_features: Tensor, logit_scale: Tensor) -> Tensor:
    """Zero-shot branch -- VitaCLIP_model.py:296-306."""
    v = video_features / video_features.norm(dim=-1, keepdim=True)
    t = text_features / text_features.norm(dim=-1, keepdim=True)
    return logit_scale.exp() * v @ t.t()


def memory_head(sd: SD, text_features: Tensor, memory: Tensor, n_cls: int) -> Tensor:
    """Support-memory branch -- VitaCLIP_model.py:368-383 (no sigmoid loss)."""
    mem = memory.mean(dim=1)
    cols = []
    for c in range(n_cls):
        tf = F.linear(torch.tanh(F.linear(text_features[c], sd["tf_project.0.weight"], sd["tf_project.0.bias"])),
                      sd["tf_project.2.weight"], sd["tf_project.2.bias"])
        tf = tf / tf.norm(dim=-1, keepdim=True)
        m = F.linear(torch.tanh(F.linear(mem, sd[f"memory_project.{c}.0.weight"], sd[f"memory_project.{c}.0.bias"])),
                     sd[f"memory_project.{c}.2.weight"], sd[f"memory_project.{c}.2.bias"])
        m = m / m.norm(dim=-1, keepdim=True)
        cols.append((sd["logit_scale_mt"] * m @ tf.t()).unsqueeze(-1))
    return F.log_softmax(torch.cat(cols, dim=1), dim=-1)


def nte_head(sd: SD, summary: Tensor, video_nte: Tensor) -> Tensor:
    """Video <-> numeric-text-embedding branch -- VitaCLIP_model.py:311-343.

    summary: (B, D) temporal mean of the summary tokens; video_nte: (B, NUM_COMB, E).
    """
    sp = F.linear(summary, sd["sum_proj.weight"], sd["sum_proj.bias"])
    sp = sp / sp.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        valid = (video_nte.sum(dim=-1).sum(dim=-1) != 0).float()
        valid_mat = valid.unsqueeze(1) * valid.unsqueeze(0)
    nte = video_nte / video_nte.norm(dim=-1, keepdim=True)
    sim = torch.bmm(sp.unsqueeze(0).expand(nte.shape[1], -1, -1), nte.permute(1, 2, 0)).mean(0)
    mat = sd["logit_scale_vm"] * (sim * valid_mat)
    return F.log_softmax(mat, dim=-1) + F.log_softmax(mat, dim=-2)


# --------------------------------------------------------------------------- #
# whole path
# --------------------------------------------------------------------------- #
def vita_forward(
    sd: SD, video: Tensor, tokenized: Sequence[Tensor], *, num_frames: int, num_heads: int,
    num_layers: int, text_heads: int = 8, text_layers: int = 12, n_ctx: int = 8,
    num_global_prompts: int = 8, kapt: bool = False, cntn_embeds=None, taps: Optional[dict] = None,
    use_summary: bool = True, use_local: bool = True,
) -> Dict[str, Tensor]:
    """VitaCLIP.forward(x) with text prompt learning -- VitaCLIP_model.py:241-309."""
    with torch.no_grad():
        feat, summary = vision_encoder(
            sd, video, num_frames=num_frames, num_heads=num_heads, num_layers=num_layers,
            num_global_prompts=num_global_prompts, taps=taps, use_summary=use_summary, use_local=use_local)
        prompts = assemble_prompts(sd, tokenized, n_ctx, kapt=kapt, cntn_embeds=cntn_embeds)
        tfs = [text_encoder(sd, p, t, heads=text_heads, layers=text_layers) for p, t in zip(prompts, tokenized)]
        logits, text_features = logit_head(feat, tfs, sd["logit_scale"], sd.get("logit_bias"))
    return {"logits": logits, "video_features": feat, "summary": summary,
            "text_features": text_features, "class_text_features": tfs}


def rel_err(a: Tensor, ref: Tensor) -> float:
    """SURVEY.md §8(d) parity metric: max|a-ref| / max|ref|."""
    return float((a.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))
