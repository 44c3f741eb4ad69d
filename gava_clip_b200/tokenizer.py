"""CLIP byte-pair tokenizer (host-side, construction time only).

Same token ids as the reference's `SimpleTokenizer` / `tokenize`
(training/VitaCLIP_text_encoder_utils.py:62-132, training/VitaCLIP_text_encoder.py:27-65) for the same
merge table.  The merge table itself (OpenAI CLIP's `bpe_simple_vocab_16e6.txt.gz`) is third-party data
that is NOT vendored here; it is looked up, in order, at
    $GAVA_BPE_VOCAB, <this package>/bpe_simple_vocab_16e6.txt.gz, ./training/…, ./…, /root/reference/training/…
Prompts whose ids were pre-registered with `register_tokens()` (e.g. from a fixture) need no merge table.
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from pathlib import Path
from typing import Dict, Iterable, List, Sequence, Tuple, Union

import torch

VOCAB_FILE = "bpe_simple_vocab_16e6.txt.gz"
SOT, EOT = "<|startoftext|>", "<|endoftext|>"
SOT_ID, EOT_ID = 49406, 49407

_pretokenized: Dict[str, List[int]] = {}


def register_tokens(table: Dict[str, Sequence[int]]) -> None:
    """text -> full id sequence (SOT ... EOT, unpadded); lets fixtures stand in for the merge table."""
    for k, v in table.items():
        _pretokenized[k] = [int(i) for i in v]


def find_vocab() -> Path:
    cands = [os.environ.get("GAVA_BPE_VOCAB"), Path(__file__).resolve().parent / VOCAB_FILE,
             Path.cwd() / "training" / VOCAB_FILE, Path.cwd() / VOCAB_FILE,
             Path("/root/reference/training") / VOCAB_FILE]
    for c in cands:
        if c and Path(c).is_file():
            return Path(c)
    raise FileNotFoundError(
        f"CLIP merge table {VOCAB_FILE} not found; set GAVA_BPE_VOCAB or call tokenizer.register_tokens()")


def _byte_alphabet() -> Dict[int, str]:
    """byte -> printable unicode stand-in (GPT-2/CLIP reversible byte encoding)."""
    keep = [*range(ord("!"), ord("~") + 1), *range(ord("¡"), ord("¬") + 1), *range(ord("®"), ord("ÿ") + 1)]
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


class SimpleTokenizer:
    def __init__(self, bpe_path: Union[str, Path, None] = None):
        import regex
        path = Path(bpe_path) if bpe_path else find_vocab()
        lines = gzip.open(path).read().decode("utf-8").split("\n")
        merges = [tuple(l.split()) for l in lines[1:49152 - 256 - 2 + 1]]
        self.byte_encoder = _byte_alphabet()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        # vocabulary order: single bytes in printable-first order, their end-of-word forms, merges, specials
        alphabet = sorted(self.byte_encoder.values(), key=lambda c: (ord(c) >= 256, ord(c)))
        vocab = alphabet + [c + "</w>" for c in alphabet] + ["".join(m) for m in merges] + [SOT, EOT]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.ranks = {m: i for i, m in enumerate(merges)}
        self.cache = {SOT: SOT, EOT: EOT}
        self.splitter = regex.compile(
            r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
            regex.IGNORECASE)

    def _merge_word(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        parts: List[str] = list(token[:-1]) + [token[-1] + "</w>"]
        while len(parts) > 1:
            best, best_rank = None, None
            for pair in zip(parts, parts[1:]):
                r = self.ranks.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = pair, r
            if best is None:
                break
            merged, i = [], 0
            while i < len(parts):
                if i + 1 < len(parts) and (parts[i], parts[i + 1]) == best:
                    merged.append(parts[i] + parts[i + 1])
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        out = " ".join(parts)
        self.cache[token] = out
        return out

    @staticmethod
    def _clean(text: str) -> str:
        try:  # the reference runs ftfy.fix_text first; identity for the ASCII class names it ships
            import ftfy
            text = ftfy.fix_text(text)
        except ImportError:
            pass
        text = html.unescape(html.unescape(text)).strip()
        return " ".join(text.split()).strip().lower()

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for tok in self.splitter.findall(self._clean(text)):
            tok = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[p] for p in self._merge_word(tok).split(" "))
        return ids

    def decode(self, ids: Iterable[int]) -> str:
        text = "".join(self.decoder[int(i)] for i in ids)
        return bytearray(self.byte_decoder[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")


@lru_cache(maxsize=1)
def default_tokenizer() -> SimpleTokenizer:
    """One shared instance (the reference re-reads the gz table per prompt: 49 s for 400 classes)."""
    return SimpleTokenizer()


def encode_with_specials(text: str) -> List[int]:
    if text in _pretokenized:
        return list(_pretokenized[text])
    tk = default_tokenizer()
    return [SOT_ID] + tk.encode(text) + [EOT_ID]


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False) -> torch.Tensor:
    """int32 [n, context_length], zero padded: SOT + ids + EOT (text_encoder.py:27-65)."""
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.int)
    for i, t in enumerate(texts):
        ids = encode_with_specials(t)
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = EOT_ID
        out[i, : len(ids)] = torch.tensor(ids, dtype=torch.int)
    return out
