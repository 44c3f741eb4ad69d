"""Drop-in `VitaCLIP` (GaVA-CLIP) model: same constructor, forward signature, attributes and state_dict
layout as the reference training/VitaCLIP_model.py:22-401, with the video tower and the logit head
executed by sm_100a kernels through the C ABI.

    forward(x, memory=None, video_nte=None, desc_wise=False) -> (logits, logits_mt, logits_vm)

Hot path per call: TowerEngine.encode_frames (vision_encoder.py) -> gava_head_fwd (temporal mean, L2 norm,
per-class mean cosine * exp(logit_scale) + bias).  The class text features are video independent and come
from the cached `TextTower` (text_encoder.py).  The support-memory and NTE branches (reference :311-398) are
tiny (B x n_cls) auxiliary heads used only by the fine-tune recipe; they are kept in PyTorch.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .text_encoder import CLIPTextEncoder, TextPromptLearner, TextTower
from .vision_encoder import CLIPVisionEncoder

Tensor = torch.Tensor
NUM_COMB = 70  # video_dataset/dataset.py:19 (number of gait-parameter combinations per NTE)


class VitaCLIP(nn.Module):
    def __init__(
        self,
        backbone_path: str = '',
        input_size: Tuple[int, int] = (224, 224),
        num_frames: int = 16,
        use_fp16: bool = False,
        cls_type: str = 'updrs',
        num_classes: int = 4,
        feature_dim: int = 768,
        patch_size: Tuple[int, int] = (16, 16),
        num_heads: int = 12,
        num_layers: int = 12,
        mlp_factor: float = 4.0,
        embed_dim: int = 512,
        use_summary_token: bool = False,
        use_local_prompts: bool = False,
        use_global_prompts: bool = False,
        num_global_prompts: int = 8,
        use_text_prompt_learning: bool = False,
        text_context_length: int = 77,
        text_vocab_size: int = 49408,
        text_transformer_width: int = 512,
        text_transformer_heads: int = 8,
        text_transformer_layers: int = 12,
        text_num_prompts: int = 8,
        text_prompt_pos: str = 'end',
        text_prompt_init: str = '',
        text_prompt_CSC: bool = False,
        text_prompt_classes_path: str = '',
        knowledge_version: List[str] = ['v0'],
        use_descriptor: bool = False,
        token_wise_mlp: bool = False,
        zeroshot_evaluation: bool = False,
        zeroshot_text_features_path: str = '',
        use_support_memory: bool = False,
        detach_features: bool = False,
        memory_batch_size: int = 64,
        add_nte: bool = False,
        use_sigmoid_loss: bool = False,
    ):
        super().__init__()
        self.fp16 = use_fp16
        self.num_frames = num_frames
        self.num_classes = num_classes
        self.text_context_length = text_context_length
        self.text_transformer_width = text_transformer_width
        self.use_summary_token = use_summary_token
        self.use_sigmoid_loss = use_sigmoid_loss
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.logit_bias = None
        self.zeroshot_evaluation = zeroshot_evaluation
        if zeroshot_evaluation:
            self.text_features = torch.load(zeroshot_text_features_path, map_location='cpu')['text_features']

        self.visual = CLIPVisionEncoder(
            input_size=input_size, num_frames=num_frames, feature_dim=feature_dim, patch_size=patch_size,
            num_heads=num_heads, num_layers=num_layers, mlp_factor=mlp_factor, embed_dim=embed_dim,
            use_summary_token=use_summary_token, use_local_prompts=use_local_prompts,
            use_global_prompts=use_global_prompts, num_global_prompts=num_global_prompts)

        self.use_text_prompt_learning = use_text_prompt_learning
        if use_text_prompt_learning:
            self.textual = CLIPTextEncoder(
                embed_dim=embed_dim, context_length=text_context_length, vocab_size=text_vocab_size,
                transformer_width=text_transformer_width, transformer_heads=text_transformer_heads,
                transformer_layers=text_transformer_layers)

        if backbone_path:
            self.load_state_dict(torch.load(backbone_path), strict=False)

        self.use_support_memory = use_support_memory
        self.memoty_batch_size = memory_batch_size   # (sic) attribute name of the reference, :141
        self.detach_features = detach_features
        self.add_nte = add_nte

        if add_nte:
            self.sum_proj = nn.Linear(feature_dim, embed_dim)
            self.logit_scale_vm = nn.Parameter(torch.ones([]) * (np.log(10.) if use_sigmoid_loss else 100.))

        if use_support_memory:
            def _proj():
                return nn.Sequential(nn.Linear(embed_dim, embed_dim // 4), nn.Tanh(),
                                     nn.Linear(embed_dim // 4, embed_dim // 8))
            self.tf_project = _proj()
            self.memory_project = nn.ModuleList([_proj() for _ in range(num_classes)])
            if use_sigmoid_loss:
                self.logit_scale_mt = nn.Parameter(torch.ones([]) * np.log(10.))
                self.logit_bias_mt = nn.Parameter(torch.ones([]) * -10.)
            else:
                self.logit_scale_mt = nn.Parameter(torch.ones([]) * 100.)
                self.logit_bias_mt = None

        if use_sigmoid_loss:
            self.logit_scale = nn.Parameter(torch.ones([]) * np.log(np.log(10.)))
            self.logit_bias = nn.Parameter(torch.ones([]) * -10.)

        self._text_tower: Optional[TextTower] = None
        if use_text_prompt_learning:
            with open(text_prompt_classes_path, 'r') as f:
                classes = [c for c in f.read().strip().split('\n') if c[0] != '*']
            self.prompt_learner = TextPromptLearner(
                classnames=classes, text_model=self.textual, num_prompts=text_num_prompts,
                prompts_init=text_prompt_init, CSC=text_prompt_CSC, ctx_pos=text_prompt_pos, cls_type=cls_type,
                knowledge_version=knowledge_version, use_descriptor=use_descriptor, token_wise_mlp=token_wise_mlp)
            self.tokenized_prompts = self.prompt_learner.tokenized_prompts
            self._text_tower = TextTower(self.textual, self.prompt_learner)

        self._freeze_visual_except_prompts_time_embed()
        if use_text_prompt_learning:
            self._freeze_textual()
        self._scale_cache = (None, 0.0)
        self._zs_cache = None
        self.text_side_stream = os.environ.get("GAVA_TEXT_SIDE_STREAM", "0") == "1"   # fine-tune step: text side on a side stream (measured neutral)

    # ------------------------------------------------------------------ freezing (reference :230-239)
    def _freeze_visual_except_prompts_time_embed(self):
        for name, p in self.visual.named_parameters():
            if not any(k in name for k in ('summary', 'local', 'global', 'time_embed')):
                p.requires_grad = False

    def freeze_visual(self) -> None:
        """Freeze the WHOLE video tower, prompts included (not a reference method): text-side-only fine-tuning; the tower
        then runs its inference schedule (CUDA-graph replay) inside the training step."""
        for p in self.visual.parameters():
            p.requires_grad = False

    def _freeze_textual(self):
        for p in self.textual.parameters():
            p.requires_grad = False

    # ------------------------------------------------------------------ helpers
    def _scale(self) -> float:
        """exp(logit_scale) as a host float, re-read only when the parameter changes (no per-call sync)."""
        key = (self.logit_scale.data_ptr(), self.logit_scale._version)
        if self._scale_cache[0] != key:
            self._scale_cache = (key, float(self.logit_scale.detach().exp()))
        return self._scale_cache[1]

    def _zeroshot_operand(self, device):
        tf = self.text_features
        key = (tf.data_ptr(), tf._version, str(device))
        if self._zs_cache is None or self._zs_cache[0] != key:
            t = tf.to(device, torch.float32)
            t = (t / t.norm(dim=-1, keepdim=True)).contiguous()
            offs = torch.arange(t.shape[0] + 1, dtype=torch.int32, device=device)
            self._zs_cache = (key, t, offs)
        return self._zs_cache[1], self._zs_cache[2]

    def _head(self, ff: Tensor, st: Tensor, t_hat: Tensor, offs: Tensor, B: int, T: int):
        dev = ff.device
        n_out = offs.numel() - 1
        vf = torch.empty(B, ff.shape[1], device=dev, dtype=torch.float32)
        so = torch.empty(B, st.shape[1], device=dev, dtype=torch.float32) if st is not None else None   # no summary token
        logits = torch.empty(B, n_out, device=dev, dtype=torch.float32)
        if ff.shape[1] != t_hat.shape[1]:      # frame_feat rows are padded to the GEMM N tile
            pad = torch.zeros(t_hat.shape[0], ff.shape[1], device=dev, dtype=torch.float32)
            pad[:, : t_hat.shape[1]] = t_hat
            t_hat = pad
        ops.head(ff, st, t_hat, offs, vf, so, logits, B=B, T=T, scale=self._scale(), logit_bias=0.0)
        return logits, vf, so

    # ------------------------------------------------------------------ fine-tune mode
    def _text_side_needs_grad(self) -> bool:
        if not (self.training and torch.is_grad_enabled()):
            return False
        if not self.use_text_prompt_learning:
            return self.logit_scale.requires_grad
        return self.logit_scale.requires_grad or any(p.requires_grad for p in self.prompt_learner.parameters())

    def _text_features_on_side_stream(self, dev):
        main = torch.cuda.current_stream(dev)
        side = getattr(self, "_text_stream", None)
        if side is None or side.device != dev:
            side = self._text_stream = torch.cuda.Stream(dev)
        side.wait_stream(main)                       # parameters updated by the optimizer on the main stream
        with torch.cuda.stream(side):
            out = self._text_tower.features(dev, need_grad=True)
        done = torch.cuda.Event()
        done.record(side)
        for t in (out[0], out[2]):
            t.record_stream(main)                    # consumed by the head on the main stream
        return out, done

    def _forward_finetune(self, ff: Tensor, st: Tensor, B: int, T: int, memory, video_nte, text_ahead=None):
        """Training-mode forward (train.py:441-443).  The video tower ran on the sm_100a kernels; when any of its trainable
        parameters requires grad, `ff`/`st` come from train_engine.TowerFunction and loss.backward() drives the backward
        kernels.  The head -- temporal mean, cosine logits (VitaCLIP_model.py:248-293) -- and the text side (prompt learner
        / KAPT projector -> text transformer, R description rows) are a few MFLOP..GFLOP and run under PyTorch autograd,
        which yields the gradients of prompt_learner.*, logit_scale and the memory / NTE heads."""
        dev = ff.device
        E = self.visual.embed_dim
        vf = ff[:, :E].reshape(B, T, E).mean(dim=1)
        summary = st.reshape(B, T, -1).mean(dim=1) if st is not None else None
        v_hat = vf / vf.norm(dim=-1, keepdim=True)
        scale = self.logit_scale.exp()
        if self.use_text_prompt_learning:
            if text_ahead is not None:
                (t_hat, offs, text_features, counts), done = text_ahead
                torch.cuda.current_stream(dev).wait_event(done)
            else:
                t_hat, offs, text_features, counts = self._text_tower.features(dev, need_grad=self._text_side_needs_grad())
            logits = torch.cat([(scale * v_hat @ t.t()).mean(-1, keepdim=True) for t in t_hat.split(counts, dim=0)], dim=-1)
            self.text_features = text_features
        else:
            t_hat, _ = self._zeroshot_operand(dev)
            logits = scale * v_hat @ t_hat.t()
        self.video_features, self.summary = vf, summary
        if self.logit_bias is not None:
            logits = logits + self.logit_bias
        return logits, self._memory_logits(memory), self._nte_logits(summary, video_nte)

    # ------------------------------------------------------------------ forward (reference :241-401)
    def forward(self, x: Tensor, memory=None, video_nte=None, desc_wise=False):
        """x: fp32 clip (B,3,T,H,W) as in the reference, or -- after self.visual.set_input_norm(mean, std) -- the
        decoder's uint8 frames (B,T,H,W,3), normalised on the fly inside the patch gather."""
        if x.is_cuda and x.device.index != torch.cuda.current_device():
            with torch.cuda.device(x.device):        # the kernels launch on the current device / stream
                return self._forward(x, memory, video_nte, desc_wise)
        return self._forward(x, memory, video_nte, desc_wise)

    def _forward(self, x: Tensor, memory=None, video_nte=None, desc_wise=False):
        text_ahead = None
        if self.use_text_prompt_learning and x.is_cuda and self._text_side_needs_grad() and self.text_side_stream:
            # fine-tune step: the text side (a few hundred small library kernels, launch-bound) is issued FIRST on a side
            # stream so that it runs underneath the video tower instead of in front of it; autograd replays its backward
            # on the same side stream, i.e. underneath the tower's backward
            text_ahead = self._text_features_on_side_stream(x.device)
        ff, st, B, T = self.visual.frames_forward(x)
        dev = ff.device
        if self.training and torch.is_grad_enabled() and (ff.requires_grad or self._text_side_needs_grad()):
            assert not desc_wise
            return self._forward_finetune(ff, st, B, T, memory, video_nte, text_ahead)
        if self.use_text_prompt_learning:
            t_hat, offs, text_features, counts = self._text_tower.features(dev)
            if desc_wise:
                assert self.training is False
                row_offs = torch.arange(t_hat.shape[0] + 1, dtype=torch.int32, device=dev)
                rows, vf, summary = self._head(ff, st, t_hat, row_offs, B, T)
                logits = list(rows.split(counts, dim=1))
                # reference quirk (:272-276,293): desc_wise leaves the LAST class's per-description features here
                text_features = t_hat[t_hat.shape[0] - counts[-1]:]
            else:
                logits, vf, summary = self._head(ff, st, t_hat, offs, B, T)
            self.text_features = text_features
        else:
            t_hat, offs = self._zeroshot_operand(dev)
            logits, vf, summary = self._head(ff, st, t_hat, offs, B, T)
        self.video_features = vf[:, : self.visual.embed_dim]     # un-normalised (what self.visual(x) returns)
        self.summary = summary

        if self.logit_bias is not None:
            if isinstance(logits, list):
                logits = [l + self.logit_bias for l in logits]
            else:
                logits = logits + self.logit_bias

        return logits, self._memory_logits(memory), self._nte_logits(summary, video_nte)

    def _nte_logits(self, summary: Tensor, video_nte):
        """video <-> numeric-text-embedding contrastive head (reference :311-343); tiny, PyTorch."""
        if not (self.add_nte and video_nte is not None):
            return None
        sp = self.sum_proj(summary)
        sp = sp / sp.norm(dim=-1, keepdim=True)
        with torch.no_grad():
            valid = (video_nte.sum(dim=-1).sum(dim=-1) != 0).float()
            valid_mat = valid.unsqueeze(1) * valid.unsqueeze(0)
        nte = video_nte / video_nte.norm(dim=-1, keepdim=True)
        sim = torch.bmm(sp.unsqueeze(0).expand(NUM_COMB, -1, -1), nte.permute(1, 2, 0)).mean(0)
        mat = self.logit_scale_vm * (sim * valid_mat)
        return F.log_softmax(mat, dim=-1) + F.log_softmax(mat, dim=-2)

    def _memory_logits(self, memory):
        """support-memory head (reference :347-391); tiny, PyTorch."""
        if not (self.use_support_memory and memory is not None):
            return None
        tfeat = self.text_features.detach() if self.detach_features else self.text_features
        mem = memory.mean(dim=1)
        cols = []
        for cid, mproj in enumerate(self.memory_project):
            tf = self.tf_project(tfeat[cid])
            tf = tf / tf.norm(dim=-1, keepdim=True)
            m = mproj(mem)
            m = m / m.norm(dim=-1, keepdim=True)
            cols.append((self.logit_scale_mt * m @ tf.t()).unsqueeze(-1))
        logits_mt = F.log_softmax(torch.cat(cols, dim=1), dim=-1)
        if self.logit_bias_mt is not None:
            logits_mt = logits_mt + self.logit_bias_mt
        return logits_mt
