"""Shape configuration of the accelerated path (video tower + logit head).

Field names follow the reference constructor kwargs
(`training/VitaCLIP_model.py:24-74`, `training/VitaCLIP_vision_encoder.py:20-40`).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Tuple


@dataclass(frozen=True)
class TowerConfig:
    input_size: Tuple[int, int] = (224, 224)
    num_frames: int = 8
    feature_dim: int = 768
    patch_size: Tuple[int, int] = (16, 16)
    num_heads: int = 12
    num_layers: int = 12
    mlp_factor: float = 4.0
    embed_dim: int = 512
    num_global_prompts: int = 8

    @property
    def grid(self) -> Tuple[int, int]:
        return (self.input_size[0] // self.patch_size[0], self.input_size[1] // self.patch_size[1])

    @property
    def num_patches(self) -> int:
        g = self.grid
        return g[0] * g[1]

    @property
    def tokens_per_frame(self) -> int:
        """live rows per frame: cls + patches (N0 in SURVEY.md)."""
        return self.num_patches + 1

    @property
    def mlp_dim(self) -> int:
        return round(self.mlp_factor * self.feature_dim)

    @property
    def head_dim(self) -> int:
        return self.feature_dim // self.num_heads

    @property
    def attn_keys(self) -> int:
        """S = N0 + G + T + 1 keys seen by every live query row."""
        return self.tokens_per_frame + self.num_global_prompts + self.num_frames + 1

    @property
    def patch_k(self) -> int:
        return 3 * self.patch_size[0] * self.patch_size[1]

    def flops_per_clip(self, n_text_rows: int = 0) -> float:
        """Reference-equivalent dense FLOPs per clip (SURVEY.md §8(d) formula)."""
        T, D, L, E = self.num_frames, self.feature_dim, self.num_layers, self.embed_dim
        Np, N0, G = self.num_patches, self.tokens_per_frame, self.num_global_prompts
        S = N0 + G + 1 + T
        M = N0 + G
        Dm = self.mlp_dim
        per_frame = 2 * Np * D * self.patch_k + L * (8 * S * D * D + 4 * S * S * D + 4 * M * D * Dm
                                                     + 10 * D * D + 4 * T * D) + 2 * D * E
        return float(T * per_frame + 2 * E * n_text_rows)

    def as_dict(self):
        return asdict(self)


VIT_B16_T8 = TowerConfig(num_frames=8)
VIT_B16_T16 = TowerConfig(num_frames=16)
VIT_L14_T16 = TowerConfig(num_frames=16, feature_dim=1024, patch_size=(14, 14), num_heads=16,
                          num_layers=24, embed_dim=768)
