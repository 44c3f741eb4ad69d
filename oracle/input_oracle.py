"""CPU oracle of the evaluation input stage.  TEST INFRASTRUCTURE ONLY (see vita_oracle.py's header for the rules).

Plain-PyTorch restatement of the eval branch of the reference dataset, video_dataset/dataset.py:117-139 (normalise,
permute, bilinear resize) and :160-196 (temporal / spatial crops).  Pinned against the reference's own
VideoDataset.__getitem__ / crop methods run in the build container on synthetic frames served through a fake `av`
module (oracle/make_golden_input.py -> tests/golden/input_stage.pt; checked by tests/test_oracle_golden.py).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# case name -> (T0, H0, W0, spatial_size, num_frames, sampling_rate, num_spatial_views, num_temporal_views)
INPUT_CASES = {
    "landscape": (11, 40, 56, 32, 4, 2, 1, 1),
    "portrait": (9, 50, 36, 32, 4, 1, 1, 1),
    "short_video": (3, 33, 47, 32, 4, 2, 1, 1),          # fewer frames than the window: last frame repeated
    "views_3x2": (12, 36, 60, 32, 4, 2, 3, 2),
    "upscale": (6, 20, 28, 32, 4, 1, 1, 3),              # source smaller than the target size
}
MEAN = (0.48145466, 0.4578275, 0.40821073)
STD = (0.26862954, 0.26130258, 0.27577711)


def make_frames(name: str) -> Tensor:
    T0, H0, W0 = INPUT_CASES[name][:3]
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    return torch.randint(0, 256, (T0, H0, W0, 3), dtype=torch.uint8, generator=g)


def temporal_crops(frames: Tensor, num_frames: int, sampling_rate: int, n_views: int) -> List[Tensor]:
    """dataset.py:160-175; frames (C, T, H, W)."""
    seg_len = (num_frames - 1) * sampling_rate + 1
    if frames.size(1) < seg_len:
        frames = torch.cat([frames, frames[:, -1:].repeat(1, seg_len - frames.size(1), 1, 1)], dim=1)
    slide_len = frames.size(1) - seg_len
    crops = []
    for i in range(n_views):
        st = slide_len // 2 if n_views == 1 else round(slide_len / (n_views - 1) * i)
        crops.append(frames[:, st: st + num_frames * sampling_rate: sampling_rate])
    return crops


def spatial_crops(frames: Tensor, spatial_size: int, n_views: int) -> List[Tensor]:
    """dataset.py:178-196."""
    S = spatial_size
    if n_views == 1:
        h_st, w_st = (frames.size(-2) - S) // 2, (frames.size(-1) - S) // 2
        return [frames[:, :, h_st: h_st + S, w_st: w_st + S]]
    margin = max(frames.size(-2), frames.size(-1)) - S
    out = []
    for st in (0, margin // 2, margin):
        out.append(frames[:, :, st: st + S, :] if frames.size(-2) > frames.size(-1) else frames[:, :, :, st: st + S])
    return out


def eval_views(frames_u8: Tensor, spatial_size: int, num_frames: int, sampling_rate: int, n_spatial: int, n_temporal: int,
               mean: Sequence[float] = MEAN, std: Sequence[float] = STD) -> List[Tensor]:
    """dataset.py:117-139: uint8 (T0, H0, W0, 3) -> list of fp32 (3, T, S, S) views, spatial-major."""
    frames = frames_u8.float() / 255.
    frames = (frames - torch.tensor(mean)) / torch.tensor(std)
    frames = frames.permute(3, 0, 1, 2)
    S = spatial_size
    if frames.size(-2) < frames.size(-1):
        new_width, new_height = frames.size(-1) * S // frames.size(-2), S
    else:
        new_height, new_width = frames.size(-2) * S // frames.size(-1), S
    frames = F.interpolate(frames, size=(new_height, new_width), mode="bilinear", align_corners=False)
    views = spatial_crops(frames, S, n_spatial)
    return sum([temporal_crops(x, num_frames, sampling_rate, n_temporal) for x in views], [])
