"""Tensor-level wrappers over the C ABI (include/gava_sm100.h).

PyTorch is used only for device memory and streams: every function takes CUDA tensors, checks
dtype/layout, and launches the sm_100a kernel on the current stream.  Outputs are preallocated by the
caller (or allocated here when `out=None`).  No function has a fallback implementation.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32, EPI_BIAS_RESID_STATS_F32,
                   EPI_DGELU_BF16, EPI_LNFOLD_BF16, EPI_PATCH_F32, GemmEpilogue)

Tensor = torch.Tensor


_launches = 0


def launch_count() -> int:
    """number of sm_100a kernel launches issued through this module so far (bench.py's `gpu_launches`)."""
    return _launches


def _stream() -> int:
    global _launches
    _launches += 1          # every wrapper below fetches the stream exactly once, right before its one launch
    return torch.cuda.current_stream().cuda_stream


def _req(t: Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the sm_100a path has no CPU fallback)")
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the CURRENT device and stream; a tensor living elsewhere would be read through a wrong
        # context (or fault).  The towers wrap their schedules in torch.cuda.device(...), direct callers must too.
        raise RuntimeError(f"{name}: tensor is on {t.device} but the current CUDA device is {torch.cuda.current_device()} "
                           "(use `with torch.cuda.device(t.device):`)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def gemm(a: Tensor, w: Tensor, out: Tensor, *, mode: int, bias: Optional[Tensor] = None,
         resid: Optional[Tensor] = None, pos: Optional[Tensor] = None, time_embed: Optional[Tensor] = None,
         group: int = 0, group_stride: int = 0, row_offset: int = 0, T: int = 0, aux: Optional[Tensor] = None,
         aux_out: Optional[Tensor] = None, max_ctas: int = 0, copy_out: Optional[Tensor] = None,
         stats_out: Optional[Tensor] = None, stats_in: Optional[Tensor] = None, stats_row_step: int = 1,
         colsum: Optional[Tensor] = None, eps: float = 1e-5) -> Tensor:
    """out = epilogue(a @ w.T).  a: bf16 [M,K] (row stride free), w: bf16 [N,K], fp32 bias [N].

    LayerNorm fold: mode EPI_BIAS_RESID_STATS_F32 (producer) also writes copy_out = bf16(out) and stats_out
    [M, gemm_stats_slots(M, N), 2]; mode EPI_LNFOLD_BF16 (consumer) takes that copy as `a`, w = W * ln.weight, bias =
    b + W @ ln.bias, colsum = w.float().sum(1) and stats_in (row r of `a` <-> stats row r * stats_row_step)."""
    _req(a, torch.bfloat16, "gemm a"); _req(w, torch.bfloat16, "gemm w")
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K, (a.shape, w.shape)
    want = torch.bfloat16 if mode in (EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_DGELU_BF16, EPI_LNFOLD_BF16) else torch.float32
    _req(out, want, "gemm out")
    assert out.dim() == 2 and out.stride(1) == 1 and out.shape[1] >= N
    if bias is not None:
        _req(bias, torch.float32, "gemm bias"); assert bias.numel() == N and bias.is_contiguous()
    epi = GemmEpilogue()
    epi.mode = mode
    epi.bias = _ptr(bias)
    epi.out = out.data_ptr()
    epi.ldo = out.stride(0)
    epi.max_ctas = max_ctas
    if mode in (EPI_BIAS_RESID_F32, EPI_BIAS_RESID_STATS_F32):
        _req(resid, torch.float32, "gemm resid")
        assert resid.dim() == 2 and resid.stride(1) == 1 and resid.shape[0] >= M
        epi.resid = resid.data_ptr(); epi.ldr = resid.stride(0)
        assert out.shape[0] >= M
        if aux is not None:      # second, bf16 addend: out = resid + aux + a @ w.T + bias
            _req(aux, torch.bfloat16, "gemm aux")
            assert aux.dim() == 2 and aux.stride(1) == 1 and aux.shape[0] >= M and aux.shape[1] >= N and aux.stride(0) % 4 == 0
            epi.aux = aux.data_ptr(); epi.ld_aux = aux.stride(0)
        if mode == EPI_BIAS_RESID_STATS_F32:
            _req(copy_out, torch.bfloat16, "gemm copy_out"); _req(stats_out, torch.float32, "gemm stats_out")
            ns = gemm_stats_slots(M, N)
            assert copy_out.dim() == 2 and copy_out.stride(1) == 1 and copy_out.shape[0] >= M and copy_out.shape[1] >= N
            assert stats_out.is_contiguous() and stats_out.shape == (M, ns, 2), (stats_out.shape, (M, ns, 2))
            epi.copy_out = copy_out.data_ptr(); epi.ld_copy = copy_out.stride(0)
            epi.stats_out = stats_out.data_ptr(); epi.n_slots = ns
    elif mode == EPI_LNFOLD_BF16:
        _req(stats_in, torch.float32, "gemm stats_in")
        assert stats_in.is_contiguous() and stats_in.dim() == 3 and stats_in.shape[2] == 2
        assert stats_in.shape[0] >= (M - 1) * stats_row_step + 1
        if colsum is not None:       # None: the rows of w are centred (sum_k w[n, k] == 0), the mean term vanishes
            _req(colsum, torch.float32, "gemm colsum"); assert colsum.is_contiguous() and colsum.numel() == N
        assert bias is not None and out.shape[0] >= M
        epi.stats_in = stats_in.data_ptr(); epi.n_slots = stats_in.shape[1]
        epi.ld_stats = 2 * stats_in.shape[1] * stats_row_step
        epi.colsum = _ptr(colsum); epi.eps = eps
    elif mode == EPI_PATCH_F32:
        _req(pos, torch.float32, "gemm pos"); _req(time_embed, torch.float32, "gemm time_embed")
        assert pos.is_contiguous() and time_embed.is_contiguous()
        assert pos.shape == (group + 1, N) and time_embed.shape == (T, N)
        assert M % group == 0 and out.shape[0] >= (M // group) * group_stride
        epi.pos = pos.data_ptr(); epi.time_embed = time_embed.data_ptr()
        epi.group, epi.group_stride, epi.row_offset, epi.T = group, group_stride, row_offset, T
    elif mode == EPI_DGELU_BF16:
        _req(aux, torch.bfloat16, "gemm aux")
        assert aux.dim() == 2 and aux.stride(1) == 1 and aux.shape[0] >= M and aux.shape[1] >= N and out.shape[0] >= M
        epi.aux = aux.data_ptr(); epi.ld_aux = aux.stride(0)
    else:
        assert out.shape[0] >= M
    if aux_out is not None:      # GELU epilogue of the training forward: keep the pre-activations too
        assert mode == EPI_BIAS_GELU_BF16
        _req(aux_out, torch.bfloat16, "gemm aux_out")
        assert aux_out.dim() == 2 and aux_out.stride(1) == 1 and aux_out.shape[0] >= M and aux_out.shape[1] >= N
        epi.aux_out = aux_out.data_ptr(); epi.ld_aux = aux_out.stride(0)
    rc = _lib.load().gava_gemm_bf16_fwd(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K,
                                        C.byref(epi), _stream())
    _lib.check(rc, "gava_gemm_bf16_fwd")
    return out


def layernorm(x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], out: Tensor, *, rows: int,
              in_row_stride: int, out_row_stride: Optional[int] = None, eps: float = 1e-5,
              delta: Optional[Tensor] = None) -> Tensor:
    """LayerNorm (or plain cast when weight is None) of `rows` rows of width D = weight/out last dim.

    Row r is read at x.data_ptr + r*in_row_stride floats: lets the caller gather e.g. the cls rows.
    """
    _req(x, torch.float32, "layernorm x")
    D = out.shape[-1]
    out_bf16 = 1 if out.dtype == torch.bfloat16 else 0
    if not out_bf16:
        _req(out, torch.float32, "layernorm out")
    if weight is not None:
        _req(weight, torch.float32, "layernorm weight"); _req(bias, torch.float32, "layernorm bias")
        assert weight.numel() == D and bias.numel() == D
    if out_row_stride is None:
        out_row_stride = D
    assert x.numel() >= (rows - 1) * in_row_stride + D and out.numel() >= (rows - 1) * out_row_stride + D
    if delta is not None:       # LayerNorm(x + delta), delta bf16 [rows, D] contiguous
        _req(delta, torch.bfloat16, "layernorm delta")
        assert delta.is_contiguous() and delta.shape[-1] == D and delta.numel() >= rows * D
        rc = _lib.load().gava_layernorm_resid_fwd(x.data_ptr(), in_row_stride, delta.data_ptr(), D, _ptr(weight), _ptr(bias),
                                                  out.data_ptr(), out_row_stride, out_bf16, rows, D, eps, _stream())
        _lib.check(rc, "gava_layernorm_resid_fwd")
        return out
    rc = _lib.load().gava_layernorm_fwd(x.data_ptr(), in_row_stride, _ptr(weight), _ptr(bias), out.data_ptr(),
                                        out_row_stride, out_bf16, rows, D, eps, _stream())
    _lib.check(rc, "gava_layernorm_fwd")
    return out


def layernorm2(x: Tensor, w: Tensor, b: Tensor, out: Tensor, w2: Tensor, b2: Tensor, out2: Tensor, *, rows: int,
               eps: float = 1e-5) -> None:
    """out (f32, may alias x) = LN(x; w, b); out2 (bf16) = LN(out; w2, b2) -- one pass over contiguous [rows, D] tensors."""
    for t in (x, w, b, out, w2, b2):
        _req(t, torch.float32, "layernorm2")
    _req(out2, torch.bfloat16, "layernorm2 out2")
    D = out.shape[-1]
    assert x.is_contiguous() and out.is_contiguous() and out2.is_contiguous() and out2.shape[-1] == D
    assert min(x.numel(), out.numel(), out2.numel()) >= rows * D and w.numel() == D and w2.numel() == D
    rc = _lib.load().gava_layernorm2_fwd(x.data_ptr(), D, w.data_ptr(), b.data_ptr(), out.data_ptr(), D, w2.data_ptr(),
                                         b2.data_ptr(), out2.data_ptr(), D, rows, D, eps, _stream())
    _lib.check(rc, "gava_layernorm2_fwd")


def patch_gather(video: Tensor, out: Tensor, patch: int) -> Tensor:
    """video fp32 [B,3,T,H,W] contiguous -> bf16 [B*T*gh*gw, Kpad] patch rows (conv-weight column order)."""
    _req(video, torch.float32, "patch_gather video"); _req(out, torch.bfloat16, "patch_gather out")
    assert video.is_contiguous() and out.is_contiguous() and video.dim() == 5 and video.shape[1] == 3
    B, _, T, H, W = video.shape
    assert out.shape[0] == B * T * (H // patch) * (W // patch)
    rc = _lib.load().gava_patch_gather_fwd(video.data_ptr(), out.data_ptr(), B, T, H, W, patch, out.shape[1], _stream())
    _lib.check(rc, "gava_patch_gather_fwd")
    return out


def cls_rows(x: Tensor, tokens_per_frame: int, cls_token: Tensor, pos_embed: Tensor, time_embed: Tensor,
             BT: int, T: int) -> Tensor:
    _req(x, torch.float32, "cls_rows x")
    for t in (cls_token, pos_embed, time_embed):
        _req(t, torch.float32, "cls_rows param"); assert t.is_contiguous()
    D = cls_token.numel()
    assert x.is_contiguous() and x.numel() >= BT * tokens_per_frame * D and time_embed.shape == (T, D)
    rc = _lib.load().gava_cls_rows_fwd(x.data_ptr(), tokens_per_frame, cls_token.data_ptr(), pos_embed.data_ptr(),
                                       time_embed.data_ptr(), BT, T, D, _stream())
    _lib.check(rc, "gava_cls_rows_fwd")
    return x


def temporal_attention(qkv: Tensor, out: Tensor, B: int, T: int, H: int) -> Tensor:
    _req(qkv, torch.bfloat16, "temporal_attention qkv"); _req(out, torch.bfloat16, "temporal_attention out")
    D = H * 64
    assert qkv.is_contiguous() and out.is_contiguous() and qkv.shape == (B * T, 3 * D) and out.shape == (B * T, D)
    rc = _lib.load().gava_temporal_attention_fwd(qkv.data_ptr(), out.data_ptr(), B, T, H, _stream())
    _lib.check(rc, "gava_temporal_attention_fwd")
    return out


def prompt_rows(global_prompts: Tensor, local_prompts: Tensor, cls_proj_out: Tensor, summary: Tensor,
                ln_w: Tensor, ln_b: Tensor, out: Tensor, B: int, T: int, eps: float = 1e-5) -> Tensor:
    for t in (global_prompts, local_prompts, cls_proj_out, summary, ln_w, ln_b):
        _req(t, torch.float32, "prompt_rows input"); assert t.is_contiguous()
    _req(out, torch.bfloat16, "prompt_rows out")
    G, D = global_prompts.shape
    assert local_prompts.shape[-2:] == (T, D) and cls_proj_out.shape == (B * T, D) and summary.shape == (B * T, D)
    assert out.is_contiguous() and out.shape == (G + 2 * B * T, D)
    rc = _lib.load().gava_prompt_rows_fwd(global_prompts.data_ptr(), local_prompts.data_ptr(), cls_proj_out.data_ptr(),
                                          summary.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), out.data_ptr(),
                                          G, B, T, D, eps, _stream())
    _lib.check(rc, "gava_prompt_rows_fwd")
    return out


def attention(qkv: Tensor, kv_aux: Tensor, out: Tensor, *, B: int, T: int, N0: int, G: int, H: int,
              n_q: Optional[int] = None, key_groups: int = 3) -> Tensor:
    """key_groups: bit 0 = local-prompt rows take part, bit 1 = the summary-token row takes part (3 = full Vita-CLIP)."""
    _req(qkv, torch.bfloat16, "attention qkv"); _req(kv_aux, torch.bfloat16, "attention kv_aux")
    _req(out, torch.bfloat16, "attention out")
    D = H * 64
    assert qkv.is_contiguous() and kv_aux.is_contiguous() and out.is_contiguous()
    assert qkv.shape == (B * T * N0, 3 * D) and kv_aux.shape == (G + 2 * B * T, 2 * D) and out.shape == (B * T * N0, D)
    rc = _lib.load().gava_attention_fwd(qkv.data_ptr(), kv_aux.data_ptr(), out.data_ptr(), B, T, N0, G, H,
                                        N0 if n_q is None else n_q, key_groups, _stream())
    _lib.check(rc, "gava_attention_fwd")
    return out


def head(frame_feat: Tensor, summary_tok: Optional[Tensor], text_hat: Tensor, class_offsets: Tensor,
         video_features: Tensor, summary_out: Optional[Tensor], logits: Tensor, *, B: int, T: int, scale: float,
         logit_bias: float = 0.0) -> Tensor:
    _req(frame_feat, torch.float32, "head frame_feat"); _req(text_hat, torch.float32, "head text_hat")
    _req(class_offsets, torch.int32, "head class_offsets")
    _req(video_features, torch.float32, "head video_features"); _req(logits, torch.float32, "head logits")
    E = frame_feat.shape[1]
    n_cls = class_offsets.numel() - 1
    D = 0
    if summary_tok is not None:
        _req(summary_tok, torch.float32, "head summary_tok"); _req(summary_out, torch.float32, "head summary_out")
        D = summary_tok.shape[1]
        assert summary_tok.is_contiguous() and summary_out.is_contiguous() and summary_out.shape == (B, D)
    assert frame_feat.is_contiguous() and text_hat.is_contiguous() and text_hat.shape[1] == E
    assert frame_feat.shape[0] == B * T and video_features.shape == (B, E) and logits.shape == (B, n_cls)
    assert video_features.is_contiguous() and logits.is_contiguous()
    rc = _lib.load().gava_head_fwd(frame_feat.data_ptr(), _ptr(summary_tok), text_hat.data_ptr(),
                                   class_offsets.data_ptr(), video_features.data_ptr(), _ptr(summary_out),
                                   logits.data_ptr(), B, T, E, D, n_cls, scale, logit_bias, _stream())
    _lib.check(rc, "gava_head_fwd")
    return logits


# ----------------------------------------------------------------------------------------------------------------------
# fp32-accuracy path: hi/lo bf16 operand split + 3-term tensor-core products, fp32 glue kernels (csrc/precise.cu)
def split_bf16(x: Tensor, hi: Tensor, lo: Tensor, *, rows: int, D: int, in_row_stride: Optional[int] = None) -> None:
    """hi = bf16(x), lo = bf16(x - hi) for `rows` rows of width D; x rows at pitch in_row_stride floats."""
    _req(x, torch.float32, "split x"); _req(hi, torch.bfloat16, "split hi"); _req(lo, torch.bfloat16, "split lo")
    in_row_stride = D if in_row_stride is None else in_row_stride
    assert hi.is_contiguous() and lo.is_contiguous() and hi.numel() >= rows * D and lo.numel() >= rows * D
    assert x.numel() >= (rows - 1) * in_row_stride + D
    rc = _lib.load().gava_split_bf16_fwd(x.data_ptr(), in_row_stride, hi.data_ptr(), lo.data_ptr(), D, rows, D, _stream())
    _lib.check(rc, "gava_split_bf16_fwd")


def split3_bf16(x: Tensor, out: Tensor, *, rows: int, D: int, in_row_stride: Optional[int] = None) -> Tensor:
    """out[r] = [hi | lo | hi] of x[r] (bf16, 3D wide): the A operand of a K-concatenated 3-term product (see pack3)."""
    _req(x, torch.float32, "split3 x"); _req(out, torch.bfloat16, "split3 out")
    in_row_stride = D if in_row_stride is None else in_row_stride
    assert out.is_contiguous() and out.numel() >= rows * 3 * D and x.numel() >= (rows - 1) * in_row_stride + D
    rc = _lib.load().gava_split3_bf16_fwd(x.data_ptr(), in_row_stride, out.data_ptr(), rows, D, _stream())
    _lib.check(rc, "gava_split3_bf16_fwd")
    return out


def pack3(w32: Tensor) -> Tensor:
    """[w_hi | w_hi | w_lo] along K (bf16 [N, 3K]): the weight operand matching split3_bf16's [hi | lo | hi] rows, so that
    A' @ W'^T = a_hi w_hi + a_lo w_hi + a_hi w_lo (the a_lo w_lo term, ~2^-18 relative, is dropped as in gemm3)."""
    w32 = w32.detach().float().contiguous()
    hi = w32.to(torch.bfloat16)
    lo = (w32 - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, hi, lo], dim=1).contiguous()


def patch_gather_split(video: Tensor, hi: Tensor, lo: Tensor, patch: int) -> None:
    _req(video, torch.float32, "patch_gather_split video"); _req(hi, torch.bfloat16, "hi"); _req(lo, torch.bfloat16, "lo")
    assert video.is_contiguous() and hi.is_contiguous() and lo.is_contiguous() and hi.shape == lo.shape
    B, _, T, H, W = video.shape
    assert hi.shape[0] == B * T * (H // patch) * (W // patch)
    rc = _lib.load().gava_patch_gather_split_fwd(video.data_ptr(), hi.data_ptr(), lo.data_ptr(), B, T, H, W, patch,
                                                 hi.shape[1], _stream())
    _lib.check(rc, "gava_patch_gather_split_fwd")


def embed_assemble(tmp: Tensor, x: Tensor, pos_embed: Tensor, time_embed: Tensor, BT: int, T: int, Np: int) -> None:
    for t in (tmp, x, pos_embed, time_embed):
        _req(t, torch.float32, "embed_assemble"); assert t.is_contiguous()
    D = tmp.shape[1]
    assert tmp.shape[0] == BT * Np and x.numel() >= BT * (Np + 1) * D and pos_embed.shape == (Np + 1, D)
    rc = _lib.load().gava_embed_assemble_fwd(tmp.data_ptr(), x.data_ptr(), pos_embed.data_ptr(), time_embed.data_ptr(),
                                             BT, T, Np, D, _stream())
    _lib.check(rc, "gava_embed_assemble_fwd")


def quickgelu_f32(x: Tensor) -> Tensor:
    _req(x, torch.float32, "quickgelu x"); assert x.is_contiguous()
    rc = _lib.load().gava_quickgelu_f32_fwd(x.data_ptr(), x.numel(), _stream())
    _lib.check(rc, "gava_quickgelu_f32_fwd")
    return x


def attention_f32(qkv: Tensor, kv_aux: Optional[Tensor], out: Tensor, *, B: int, T: int, N0: int, G: int, H: int,
                  n_q: Optional[int] = None, key_groups: int = 3) -> Tensor:
    _req(qkv, torch.float32, "attention_f32 qkv"); _req(out, torch.float32, "attention_f32 out")
    D = H * 64
    assert qkv.is_contiguous() and out.is_contiguous() and qkv.shape == (B * T * N0, 3 * D) and out.shape == (B * T * N0, D)
    if kv_aux is not None:
        _req(kv_aux, torch.float32, "attention_f32 kv_aux")
        assert kv_aux.is_contiguous() and kv_aux.shape == (G + 2 * B * T, 2 * D)
    rc = _lib.load().gava_attention_f32_fwd(qkv.data_ptr(), _ptr(kv_aux), out.data_ptr(), B, T, N0, G, H,
                                            N0 if n_q is None else n_q, key_groups, _stream())
    _lib.check(rc, "gava_attention_f32_fwd")
    return out


def prompt_rows_f32(global_prompts: Tensor, local_prompts: Tensor, cls_proj_out: Tensor, summary: Tensor, ln_w: Tensor,
                    ln_b: Tensor, out: Tensor, B: int, T: int, eps: float = 1e-5) -> Tensor:
    for t in (global_prompts, local_prompts, cls_proj_out, summary, ln_w, ln_b, out):
        _req(t, torch.float32, "prompt_rows_f32"); assert t.is_contiguous()
    G, D = global_prompts.shape
    assert out.shape == (G + 2 * B * T, D)
    rc = _lib.load().gava_prompt_rows_f32_fwd(global_prompts.data_ptr(), local_prompts.data_ptr(), cls_proj_out.data_ptr(),
                                              summary.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), out.data_ptr(),
                                              G, B, T, D, eps, _stream())
    _lib.check(rc, "gava_prompt_rows_f32_fwd")
    return out


def gemm_stats_slots(M: int, N: int) -> int:
    """column slices per row of the statistics an EPI_BIAS_RESID_STATS_F32 launch of shape (M, N) writes"""
    n = int(_lib.load().gava_gemm_stats_slots(M, N))
    if n <= 0:
        raise ValueError(f"gemm_stats_slots: unsupported shape M={M} N={N}")
    return n


def gemm3(a_hi: Tensor, a_lo: Tensor, w_hi: Tensor, w_lo: Tensor, out: Tensor, *, bias: Optional[Tensor] = None,
          resid: Optional[Tensor] = None) -> Tensor:
    """fp32-accuracy product: out(f32) = [resid +] a @ w.T + bias with a = a_hi + a_lo, w = w_hi + w_lo (bf16 pairs),
    evaluated as three tensor-core GEMMs accumulating in fp32 (the a_lo @ w_lo term, ~2^-18 relative, is dropped)."""
    if resid is not None:
        gemm(a_hi, w_hi, out, mode=EPI_BIAS_RESID_F32, bias=bias, resid=resid)
    else:
        gemm(a_hi, w_hi, out, mode=EPI_BIAS_F32, bias=bias)
    gemm(a_hi, w_lo, out, mode=EPI_BIAS_RESID_F32, resid=out)
    gemm(a_lo, w_hi, out, mode=EPI_BIAS_RESID_F32, resid=out)
    return out


def causal_attention_f32(qkv: Tensor, out: Tensor, *, n_seq: int, L: int, H: int) -> Tensor:
    """causal self-attention core of the text transformer: qkv f32 [n_seq*L, 3D] -> out f32 [n_seq*L, D]."""
    _req(qkv, torch.float32, "causal_attention qkv"); _req(out, torch.float32, "causal_attention out")
    D = H * 64
    assert qkv.is_contiguous() and out.is_contiguous() and qkv.shape == (n_seq * L, 3 * D) and out.shape == (n_seq * L, D)
    rc = _lib.load().gava_causal_attention_f32_fwd(qkv.data_ptr(), out.data_ptr(), n_seq, L, H, _stream())
    _lib.check(rc, "gava_causal_attention_f32_fwd")
    return out


def patch_gather_u8(frames: Tensor, hi: Tensor, lo: Optional[Tensor], mean, std, patch: int) -> None:
    """uint8 frames [B,T,H,W,3] -> normalised bf16 patch rows (input stage fused into the gather); mean/std: 3 floats."""
    _req(frames, torch.uint8, "patch_gather_u8 frames"); _req(hi, torch.bfloat16, "patch_gather_u8 hi")
    assert frames.is_contiguous() and frames.dim() == 5 and frames.shape[-1] == 3 and hi.is_contiguous()
    B, T, H, W, _ = frames.shape
    assert hi.shape[0] == B * T * (H // patch) * (W // patch)
    if lo is not None:
        _req(lo, torch.bfloat16, "patch_gather_u8 lo"); assert lo.is_contiguous() and lo.shape == hi.shape
    m = (C.c_float * 3)(*[float(v) for v in mean]); s = (C.c_float * 3)(*[float(v) for v in std])
    rc = _lib.load().gava_patch_gather_u8_fwd(frames.data_ptr(), hi.data_ptr(), _ptr(lo), m, s, B, T, H, W, patch,
                                              hi.shape[1], _stream())
    _lib.check(rc, "gava_patch_gather_u8_fwd")


# ----------------------------------------------------------------------------------------------------------------------
# fine-tune step: backward + optimizer kernels (csrc/backward.cu, csrc/attention_bwd.cu)
def layernorm_bwd(dy: Tensor, x: Tensor, weight: Tensor, dx_out: Tensor, *, rows: int, D: int,
                  dy_stride: Optional[int] = None, x_stride: Optional[int] = None, dx_stride: Optional[int] = None,
                  add1: Optional[Tensor] = None, add1_stride: Optional[int] = None, add2: Optional[Tensor] = None,
                  add2_stride: Optional[int] = None, dx_bf16: Optional[Tensor] = None, dxb_stride: Optional[int] = None,
                  dw: Optional[Tensor] = None, db: Optional[Tensor] = None, eps: float = 1e-5) -> Tensor:
    """dx_out = [add1] + [add2] + LayerNorm-backward(dy; x, weight); optional bf16 copy and dw/db accumulation.
    Strides are row pitches in elements (default D): lets callers address e.g. the cls rows of the stream."""
    assert dy.dtype in (torch.bfloat16, torch.float32) and dy.is_cuda
    _req(x, torch.float32, "layernorm_bwd x"); _req(weight, torch.float32, "layernorm_bwd weight")
    _req(dx_out, torch.float32, "layernorm_bwd dx_out")
    for t in (add1, add2):
        if t is not None:
            _req(t, torch.float32, "layernorm_bwd add")
    if dx_bf16 is not None:
        _req(dx_bf16, torch.bfloat16, "layernorm_bwd dx_bf16")
    if dw is not None:
        _req(dw, torch.float32, "layernorm_bwd dw"); _req(db, torch.float32, "layernorm_bwd db")
        assert dw.numel() == D and db.numel() == D
    st = lambda v: D if v is None else v
    for t, sd in ((dy, st(dy_stride)), (x, st(x_stride)), (dx_out, st(dx_stride))):
        assert t.numel() >= (rows - 1) * sd + D
    rc = _lib.load().gava_layernorm_bwd(dy.data_ptr(), st(dy_stride), 1 if dy.dtype == torch.bfloat16 else 0, x.data_ptr(),
                                        st(x_stride), weight.data_ptr(), _ptr(add1), st(add1_stride), _ptr(add2),
                                        st(add2_stride), dx_out.data_ptr(), st(dx_stride), _ptr(dx_bf16), st(dxb_stride),
                                        _ptr(dw), _ptr(db), rows, D, eps, _stream())
    _lib.check(rc, "gava_layernorm_bwd")
    return dx_out


def attention_train(qkv: Tensor, kv_aux: Tensor, out: Tensor, lse: Tensor, *, B: int, T: int, N0: int, G: int,
                    H: int, n_q: Optional[int] = None, key_groups: int = 3) -> Tensor:
    """attention() over all query rows that also saves the per-row log-sum-exp (f32 [B*T, H, N0]) for attention_bwd."""
    _req(qkv, torch.bfloat16, "attention qkv"); _req(kv_aux, torch.bfloat16, "attention kv_aux")
    _req(out, torch.bfloat16, "attention out"); _req(lse, torch.float32, "attention lse")
    D = H * 64
    assert qkv.is_contiguous() and kv_aux.is_contiguous() and out.is_contiguous() and lse.is_contiguous()
    assert qkv.shape == (B * T * N0, 3 * D) and kv_aux.shape == (G + 2 * B * T, 2 * D) and out.shape == (B * T * N0, D)
    assert lse.numel() == B * T * H * N0
    rc = _lib.load().gava_attention_train_fwd(qkv.data_ptr(), kv_aux.data_ptr(), out.data_ptr(), lse.data_ptr(), B, T, N0, G,
                                              H, N0 if n_q is None else n_q, key_groups, _stream())
    _lib.check(rc, "gava_attention_train_fwd")
    return out


def attention_bwd(qkv: Tensor, kv_aux: Tensor, out: Tensor, dout: Tensor, dqkv: Tensor, dkv_aux: Tensor, *, B: int,
                  T: int, N0: int, G: int, H: int, lse: Optional[Tensor] = None, n_q: Optional[int] = None,
                  scratch: Optional[Tensor] = None, key_groups: int = 3) -> None:
    for t in (qkv, kv_aux, out, dout, dqkv):
        _req(t, torch.bfloat16, "attention_bwd"); assert t.is_contiguous()
    _req(dkv_aux, torch.float32, "attention_bwd dkv_aux")
    D = H * 64
    assert qkv.shape == (B * T * N0, 3 * D) and dqkv.shape == qkv.shape and out.shape == (B * T * N0, D)
    assert dout.shape == out.shape and kv_aux.shape == (G + 2 * B * T, 2 * D) and dkv_aux.shape == kv_aux.shape
    assert dkv_aux.is_contiguous()
    if lse is not None:
        _req(lse, torch.float32, "attention_bwd lse"); assert lse.is_contiguous() and lse.numel() == B * T * H * N0
    if scratch is not None:
        _req(scratch, torch.float32, "attention_bwd scratch")
        assert scratch.is_contiguous() and scratch.numel() >= B * T * H * (G + T + 1) * 128
    rc = _lib.load().gava_attention_bwd(qkv.data_ptr(), kv_aux.data_ptr(), out.data_ptr(), dout.data_ptr(), _ptr(lse),
                                        dqkv.data_ptr(), dkv_aux.data_ptr(), _ptr(scratch), B, T, N0, G, H,
                                        N0 if n_q is None else n_q, key_groups, _stream())
    _lib.check(rc, "gava_attention_bwd")


def temporal_attention_bwd(qkv: Tensor, dout: Tensor, dqkv: Tensor, B: int, T: int, H: int) -> Tensor:
    for t in (qkv, dout, dqkv):
        _req(t, torch.bfloat16, "temporal_attention_bwd"); assert t.is_contiguous()
    D = H * 64
    assert qkv.shape == (B * T, 3 * D) and dqkv.shape == qkv.shape and dout.shape == (B * T, D)
    rc = _lib.load().gava_temporal_attention_bwd(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), B, T, H, _stream())
    _lib.check(rc, "gava_temporal_attention_bwd")
    return dqkv


def causal_attention_bwd(qkv: Tensor, dout: Tensor, dqkv: Tensor, *, n_seq: int, L: int, H: int) -> Tensor:
    """backward of causal_attention_f32: qkv f32 [n_seq*L, 3D], dout bf16 [n_seq*L, D] -> dqkv bf16 [n_seq*L, 3D]."""
    _req(qkv, torch.float32, "causal_attention_bwd qkv"); _req(dout, torch.bfloat16, "causal_attention_bwd dout")
    _req(dqkv, torch.bfloat16, "causal_attention_bwd dqkv")
    D = H * 64
    assert qkv.is_contiguous() and dout.is_contiguous() and dqkv.is_contiguous()
    assert qkv.shape == (n_seq * L, 3 * D) and dqkv.shape == qkv.shape and dout.shape == (n_seq * L, D)
    rc = _lib.load().gava_causal_attention_bwd(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), n_seq, L, H, _stream())
    _lib.check(rc, "gava_causal_attention_bwd")
    return dqkv


def prompt_rows_bwd(d_rows_n: Tensor, global_prompts: Tensor, local_prompts: Tensor, cls_proj_out: Tensor, summary: Tensor,
                    ln_w: Tensor, d_global: Tensor, d_local_cp: Tensor, d_summary: Tensor, d_summary_ext: Optional[Tensor],
                    B: int, T: int, eps: float = 1e-5) -> None:
    _req(d_rows_n, torch.bfloat16, "prompt_rows_bwd d_rows_n")
    for t in (global_prompts, local_prompts, cls_proj_out, summary, ln_w, d_global, d_local_cp, d_summary):
        _req(t, torch.float32, "prompt_rows_bwd"); assert t.is_contiguous()
    G, D = global_prompts.shape
    assert d_rows_n.is_contiguous() and d_rows_n.shape == (G + 2 * B * T, D)
    assert d_global.shape == (G, D) and d_local_cp.shape == (B * T, D) and d_summary.shape == (B * T, D)
    if d_summary_ext is not None:
        _req(d_summary_ext, torch.float32, "prompt_rows_bwd ext"); assert d_summary_ext.is_contiguous() and d_summary_ext.shape == (B * T, D)
    rc = _lib.load().gava_prompt_rows_bwd(d_rows_n.data_ptr(), global_prompts.data_ptr(), local_prompts.data_ptr(),
                                          cls_proj_out.data_ptr(), summary.data_ptr(), ln_w.data_ptr(), d_global.data_ptr(),
                                          d_local_cp.data_ptr(), d_summary.data_ptr(), _ptr(d_summary_ext), G, B, T, D, eps,
                                          _stream())
    _lib.check(rc, "gava_prompt_rows_bwd")


def group_sum(x: Tensor, out: Tensor, *, outer: int, J: int, inner: int, ld: Optional[int] = None) -> Tensor:
    """out[j] = sum_{o<outer, n<inner} x[(o*J + j)*inner + n]  (rows of width D = out.shape[-1], pitch ld)."""
    assert x.dtype in (torch.float32, torch.bfloat16) and x.is_cuda
    _req(out, torch.float32, "group_sum out"); assert out.is_contiguous()
    D = out.shape[-1]
    ld = D if ld is None else ld
    assert out.numel() == J * D and x.numel() >= (outer * J * inner - 1) * ld + D
    rc = _lib.load().gava_group_sum(x.data_ptr(), ld, 1 if x.dtype == torch.bfloat16 else 0, out.data_ptr(), outer, J, inner,
                                    D, _stream())
    _lib.check(rc, "gava_group_sum")
    return out


def transpose_bf16(x: Tensor, out: Tensor) -> Tensor:
    """out bf16 [C, Rpad] = x[R, C]^T, zero-padded along R (x: f32 or bf16, row stride free)."""
    assert x.dtype in (torch.float32, torch.bfloat16) and x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    _req(out, torch.bfloat16, "transpose out"); assert out.is_contiguous() and out.dim() == 2
    R, Cc = x.shape
    assert out.shape[0] == Cc and out.shape[1] >= R
    rc = _lib.load().gava_transpose_bf16(x.data_ptr(), x.stride(0), 1 if x.dtype == torch.bfloat16 else 0, out.data_ptr(),
                                         R, Cc, out.shape[1], _stream())
    _lib.check(rc, "gava_transpose_bf16")
    return out


def quickgelu_bf16(a: Tensor, h: Tensor) -> Tensor:
    _req(a, torch.bfloat16, "quickgelu a"); _req(h, torch.bfloat16, "quickgelu h")
    assert a.is_contiguous() and h.is_contiguous() and a.numel() == h.numel()
    rc = _lib.load().gava_quickgelu_bf16_fwd(a.data_ptr(), h.data_ptr(), a.numel(), _stream())
    _lib.check(rc, "gava_quickgelu_bf16_fwd")
    return h


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, *, lr: float, beta1: float = 0.9, beta2: float = 0.999,
               eps: float = 1e-8, weight_decay: float = 0.01, step: int = 1, grad_scale: float = 1.0) -> None:
    for t in (p, g, m, v):
        _req(t, torch.float32, "adamw"); assert t.is_contiguous() and t.numel() == p.numel()
    rc = _lib.load().gava_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2,
                                     eps, weight_decay, step, grad_scale, _stream())
    _lib.check(rc, "gava_adamw_step")
