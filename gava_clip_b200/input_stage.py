"""Evaluation input stage on the GPU (SURVEY.md 8(f)-3 and the view handling of 8(f)-4).

Mirrors the eval branch of the reference dataset (video_dataset/dataset.py:117-139, 160-196): decoded uint8 frames
(T0, H0, W0, 3) -> /255 -> normalise -> bilinear resize (short side to `spatial_size`) -> spatial crops (1 centre crop or
3 crops along the long side) -> temporal crops (`num_temporal_views` windows of `num_frames` frames at `sampling_rate`,
the last frame repeated for short videos).  Decoding (PyAV) stays on the host; everything after it is ONE kernel launch
per view (gava_preprocess_u8_fwd) writing the fp32 (3, T, S, S) clip VitaCLIP.forward takes, so a video crosses PCIe
once, as raw uint8.

    stage = EvalInputStage(spatial_size=224, num_frames=8, mean=[...], std=[...])
    clip  = stage(frames_u8_cuda)                     # the reference's `frames[0]`: first view, (3, T, S, S)
    views = stage.views(frames_u8_cuda)               # (n_views, 3, T, S, S), spatial-major order like the reference
    scores = aggregate_views(model(views)[0])         # mean softmax score over the views of one video
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib, ops

Tensor = torch.Tensor


class EvalInputStage:
    def __init__(self, spatial_size: int = 224, num_frames: int = 8, sampling_rate: int = 1, num_spatial_views: int = 1,
                 num_temporal_views: int = 1, mean: Sequence[float] = (0.45,), std: Sequence[float] = (0.225,)):
        if num_spatial_views not in (1, 3):
            raise NotImplementedError("num_spatial_views must be 1 or 3 (dataset.py:196)")
        mean, std = [float(v) for v in mean], [float(v) for v in std]
        if len(mean) == 1:
            mean, std = mean * 3, std * 3
        self.spatial_size, self.num_frames, self.sampling_rate = spatial_size, num_frames, sampling_rate
        self.num_spatial_views, self.num_temporal_views = num_spatial_views, num_temporal_views
        self.mean, self.std = mean, std

    # ---- geometry, exactly as the reference computes it (integer arithmetic of dataset.py:124-129, 160-194)
    def resized_hw(self, H0: int, W0: int) -> Tuple[int, int]:
        S = self.spatial_size
        if H0 < W0:
            return S, W0 * S // H0
        return H0 * S // W0, S

    def spatial_offsets(self, new_h: int, new_w: int) -> List[Tuple[int, int]]:
        S = self.spatial_size
        if self.num_spatial_views == 1:
            assert min(new_h, new_w) >= S
            return [((new_h - S) // 2, (new_w - S) // 2)]
        assert min(new_h, new_w) == S
        margin = max(new_h, new_w) - S
        return [((st, 0) if new_h > new_w else (0, st)) for st in (0, margin // 2, margin)]

    def temporal_starts(self, T0: int) -> List[int]:
        seg_len = (self.num_frames - 1) * self.sampling_rate + 1
        slide_len = max(T0, seg_len) - seg_len
        n = self.num_temporal_views
        if n == 1:
            return [slide_len // 2]
        return [round(slide_len / (n - 1) * i) for i in range(n)]

    # ---- kernel launches
    def _launch(self, frames: Tensor, out: Tensor, new_hw: Tuple[int, int], off: Tuple[int, int], t_st: int) -> None:
        T0, H0, W0, _ = frames.shape
        S = self.spatial_size
        m = (C.c_float * 3)(*self.mean); s = (C.c_float * 3)(*self.std)
        rc = _lib.load().gava_preprocess_u8_fwd(frames.data_ptr(), out.data_ptr(), T0, H0, W0, new_hw[0], new_hw[1], off[0],
                                                off[1], t_st, self.sampling_rate, self.num_frames, S, S, m, s, ops._stream())
        _lib.check(rc, "gava_preprocess_u8_fwd")

    def _check(self, frames: Tensor) -> None:
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError(f"expected decoded uint8 frames (T0, H0, W0, 3), got {tuple(frames.shape)} {frames.dtype}")
        if not frames.is_cuda:
            raise RuntimeError("EvalInputStage runs on the GPU: upload the uint8 frames first (no CPU fallback)")
        if not frames.is_contiguous():
            raise ValueError("frames must be contiguous")

    def views(self, frames: Tensor, out: Optional[Tensor] = None) -> Tensor:
        """all views of one video, spatial-major then temporal (dataset.py:137-138): (n_views, 3, T, S, S) fp32."""
        self._check(frames)
        T0, H0, W0, _ = frames.shape
        new_hw = self.resized_hw(H0, W0)
        offs, starts = self.spatial_offsets(*new_hw), self.temporal_starts(T0)
        S, T = self.spatial_size, self.num_frames
        n = len(offs) * len(starts)
        if out is None:
            out = torch.empty(n, 3, T, S, S, device=frames.device, dtype=torch.float32)
        assert out.shape == (n, 3, T, S, S) and out.is_contiguous() and out.dtype == torch.float32
        i = 0
        for off in offs:
            for st in starts:
                self._launch(frames, out[i], new_hw, off, st)
                i += 1
        return out

    def __call__(self, frames: Tensor, out: Optional[Tensor] = None) -> Tensor:
        """the clip the reference's eval loader yields: the FIRST view only (`frames = frames[0]`, dataset.py:141)."""
        self._check(frames)
        T0, H0, W0, _ = frames.shape
        new_hw = self.resized_hw(H0, W0)
        S, T = self.spatial_size, self.num_frames
        if out is None:
            out = torch.empty(3, T, S, S, device=frames.device, dtype=torch.float32)
        assert out.shape == (3, T, S, S) and out.is_contiguous() and out.dtype == torch.float32
        self._launch(frames, out, new_hw, self.spatial_offsets(*new_hw)[0], self.temporal_starts(T0)[0])
        return out

    def batch(self, videos: Sequence[Tensor], out: Optional[Tensor] = None) -> Tensor:
        """first view of each video written straight into one (B, 3, T, S, S) batch (one launch per video)."""
        S, T = self.spatial_size, self.num_frames
        if out is None:
            out = torch.empty(len(videos), 3, T, S, S, device=videos[0].device, dtype=torch.float32)
        for i, v in enumerate(videos):
            self(v, out[i])
        return out


def aggregate_views(logits: Tensor, n_views: Optional[int] = None) -> Tensor:
    """Multi-view test-time aggregation (Vita-CLIP's `scores = logits.softmax(-1).mean(0)`; the reference keeps the hook
    commented out at evaluate.py:277-279 and scores the first view only): mean softmax score over the views of a video.
    logits: (n_views, n_cls) -> (n_cls,), or (n_videos * n_views, n_cls) with n_views given -> (n_videos, n_cls)."""
    scores = logits.softmax(dim=-1)
    if n_views is None:
        return scores.mean(dim=0)
    return scores.view(-1, n_views, scores.shape[-1]).mean(dim=1)
