"""Knowledge-Aware Prompt Tuning head (mirror of training/kapt_head.py:24-214).

Loads the per-class KEPLER knowledge embeddings / natural-language descriptions from
`./data/ke_<type>/` (relative to the CWD, as the reference does, kapt_head.py:60-61) and adds a zero-initialised
class-wise MLP projection of the embeddings to the learnable context tokens.

Only the variant that actually executes in the reference is implemented in forward(): continuous
embeddings (`cntn`) projected by ONE MLP per class (`uni`, class-wise).  The other projector layouts are
still constructed with the reference's parameter names so their checkpoints load, but calling them raises,
exactly where the reference itself fails (kapt_head.py:192-195 `.expand` on a list; :201 unbound `idc`).
"""
from __future__ import annotations

import os.path as osp
from typing import List

import numpy as np
import torch
import torch.nn as nn

Tensor = torch.Tensor


def _zero_linear(m: nn.Module) -> None:
    if isinstance(m, nn.Linear):
        nn.init.zeros_(m.weight)
        if m.bias is not None:
            nn.init.zeros_(m.bias)


def _mlp(inp: int, hid: int, out: int, bias: bool) -> nn.Sequential:
    return nn.Sequential(nn.Linear(inp, hid, bias=bias), nn.ReLU(inplace=True), nn.Linear(hid, out, bias=bias))


def _read_lines(path: str) -> List[str]:
    with open(path, "r") as f:
        return [line.strip() for line in f]


class ContextualPromptLearner(nn.Module):
    def __init__(self, use_cntn: bool, cntn_split: bool, uni_mlp: bool, use_disc: bool, emb_dim: int = 128,
                 out_dim: int = 512, inp_dim: int = 768, n_cls: int = 4, n_tokens: int = 16, cls_type: str = "updrs",
                 knowledge_version: List[str] = ["v0"], use_descriptor: bool = False, token_wise_mlp: bool = True,
                 class_wise_mlp: bool = True):
        super().__init__()
        self.type = cls_type.lower().split("_")[0]
        self.n_cls, self.n_tokens = n_cls, n_tokens
        self.token_wise_mlp, self.class_wise_mlp = token_wise_mlp, class_wise_mlp
        self.use_descriptor = use_descriptor
        self.cntn, self.cntn_split, self.uni_mlp, self.disc = use_cntn, cntn_split, uni_mlp, use_disc
        self.updrs_ke_dir = f"./data/ke_{self.type}"
        assert osp.isdir(self.updrs_ke_dir), f"knowledge directory {self.updrs_ke_dir} not found (relative to CWD)"
        assert not (class_wise_mlp and token_wise_mlp), "Only one of `class_wise_mlp` and `token_wise_mlp` can be enabled."
        assert len(knowledge_version) > 0 or use_descriptor, "No knowledge is specified."
        d = self.updrs_ke_dir
        cntn: List[Tensor] = []
        if use_descriptor:
            assert use_disc or cntn_split, "Descriptor is used but disc is not enabled."
            base = np.load(osp.join(d, "all.npy"))[:n_cls]
            cls_disc = [_read_lines(osp.join(d, f"descriptor_{c}.txt")) for c in range(n_cls)]
            if use_cntn:
                for c in range(n_cls):
                    if cntn_split:
                        cntn.append(torch.from_numpy(np.load(osp.join(d, f"descriptor_{c}.npy"))).float())
                    else:
                        cntn.append(torch.from_numpy(base[c]).float().unsqueeze(0).expand(len(cls_disc[c]), -1))
        else:
            if use_cntn and not cntn_split:
                # the reference reads an unbound variable here (kapt_head.py:87-88) and cannot be constructed
                raise NotImplementedError("prompts_init without 'split' is unconstructible in the reference")
            cls_disc = [[] for _ in range(n_cls)]
            per_version = []
            for kv in knowledge_version:
                if use_cntn:
                    per_version.append(torch.from_numpy(np.load(osp.join(d, f"EntityEmb_{kv}.npy"))[:n_cls]).float())
                desc = _read_lines(osp.join(d, f"simQdesc_{kv}.txt")) if use_disc else [""] * n_cls
                for c in range(n_cls):
                    cls_disc[c].append(desc[c])
            if use_cntn:
                stacked = torch.stack(per_version, dim=1)            # (n_cls, n_kv, inp_dim)
                cntn = list(stacked)
        if use_cntn:
            if token_wise_mlp:
                self.projector = nn.ModuleList([_mlp(inp_dim, emb_dim, out_dim, True) for _ in range(n_tokens)])
            elif class_wise_mlp:
                if uni_mlp:
                    self.projector = nn.ModuleList([_mlp(inp_dim, emb_dim, out_dim, False) for _ in range(n_cls)])
                else:
                    self.projector = nn.ModuleList([
                        nn.ModuleList([_mlp(inp_dim, emb_dim, out_dim, False) for _ in range(n_tokens)])
                        for _ in range(n_cls)])
            elif uni_mlp:
                self.projector = _mlp(inp_dim, emb_dim, out_dim, True)
            else:
                self.projector = nn.ModuleList([_mlp(inp_dim, emb_dim, out_dim, True) for _ in range(n_tokens)])
            self.projector.apply(_zero_linear)
        else:
            self.projector = None
        if len(cntn) > 0:
            self.cntn_embeds = cntn          # plain list (not a buffer), moved to the device lazily
        assert len(cls_disc) > 0
        self.cls_disc = cls_disc

    def forward(self, ctx_prompt: Tensor) -> List[Tensor]:
        """ctx_prompt (n_cls, n_ctx, W) -> list[n_cls] of (n_kv, n_ctx, W): ctx + MLP_c(knowledge embedding)."""
        if not (self.cntn and self.class_wise_mlp and self.uni_mlp):
            raise NotImplementedError("only the class-wise uni-MLP knowledge prompt ('cntn_split_uni[_disc]') "
                                      "executes in the reference; other layouts fail there too")
        out = []
        for c in range(self.n_cls):
            self.cntn_embeds[c] = self.cntn_embeds[c].to(ctx_prompt.device)
            e = self.projector[c](self.cntn_embeds[c]).unsqueeze(1).expand(-1, self.n_tokens, -1)
            out.append(ctx_prompt[c].unsqueeze(0) + e)
        return out
