"""Byte-level BPE tokenizer compatible with CLIP's 49 408-entry vocabulary.

Mirrors the behaviour of the reference ``clip/simple_tokenizer.py:62-132`` (``SimpleTokenizer.encode/decode``):
lower-cased, whitespace-collapsed text is split by the CLIP regex, each piece is mapped to printable byte
symbols, and merges are applied greedily by rank with ``</w>`` marking the word end.

The merge table is the public ``bpe_simple_vocab_16e6.txt.gz`` file of OpenAI CLIP.  It is third-party data and is
not stored in this repository; it is looked up, in order, at ``$AIHAB_CLIP_BPE``, next to this module, in
``~/.cache/clip`` and (build container only) in the mounted reference checkout.
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from pathlib import Path

import regex

try:  # ftfy only repairs mojibake; it is the identity on the ASCII prompts of data/templates.py
    import ftfy

    def _fix_text(s: str) -> str:
        return ftfy.fix_text(s)
except ImportError:  # pragma: no cover - depends on the image
    def _fix_text(s: str) -> str:
        return s

_VOCAB_NAME = "bpe_simple_vocab_16e6.txt.gz"
_SPLIT = regex.compile(
    r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)


def find_vocab() -> Path:
    candidates = []
    if os.environ.get("AIHAB_CLIP_BPE"):
        candidates.append(Path(os.environ["AIHAB_CLIP_BPE"]))
    candidates += [Path(__file__).resolve().parent / _VOCAB_NAME, Path.home() / ".cache" / "clip" / _VOCAB_NAME,
                   Path("/root/reference/clip") / _VOCAB_NAME]
    for c in candidates:
        if c.is_file():
            return c
    raise FileNotFoundError(
        f"CLIP BPE merge table {_VOCAB_NAME} not found (searched: {', '.join(map(str, candidates))}); "
        "set AIHAB_CLIP_BPE to its path")


@lru_cache()
def byte_symbols() -> dict:
    """byte value -> printable unicode symbol (printable latin-1 bytes map to themselves, the rest to U+0100+)."""
    keep = [b for rng in ((0x21, 0x7E), (0xA1, 0xAC), (0xAE, 0xFF)) for b in range(rng[0], rng[1] + 1)]
    table, extra = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    return table


class SimpleTokenizer:
    def __init__(self, bpe_path: str | os.PathLike | None = None):
        path = Path(bpe_path) if bpe_path is not None else find_vocab()
        sym = byte_symbols()
        # ordered as the reference builds it: first the 188 "kept" bytes in ascending order, then the remapped ones
        base = [sym[b] for b in sorted(sym, key=lambda b: (ord(sym[b]) >= 256, ord(sym[b])))]
        lines = gzip.open(path).read().decode("utf-8").split("\n")
        merges = [tuple(ln.split()) for ln in lines[1:49152 - 256 - 2 + 1]]
        vocab = base + [s + "</w>" for s in base] + ["".join(m) for m in merges]
        vocab += ["<|startoftext|>", "<|endoftext|>"]
        self.byte_encoder = sym
        self.byte_decoder = {v: k for k, v in sym.items()}
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.bpe_ranks = {m: i for i, m in enumerate(merges)}
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}

    def bpe(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        word = list(token[:-1]) + [token[-1] + "</w>"]
        while len(word) > 1:
            ranked = [(self.bpe_ranks.get((a, b), float("inf")), i) for i, (a, b) in enumerate(zip(word, word[1:]))]
            best_rank, _ = min(ranked)
            if best_rank == float("inf"):
                break
            first, second = next((a, b) for (a, b) in zip(word, word[1:]) if self.bpe_ranks.get((a, b)) == best_rank)
            merged, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    merged.append(first + second)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = merged
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text: str) -> list:
        text = html.unescape(html.unescape(_fix_text(text))).strip()
        text = regex.sub(r"\s+", " ", text).strip().lower()
        ids = []
        for piece in _SPLIT.findall(text):
            piece = "".join(self.byte_encoder[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self.bpe(piece).split(" "))
        return ids

    def decode(self, tokens) -> str:
        text = "".join(self.decoder[int(t)] for t in tokens)
        raw = bytearray(self.byte_decoder[c] for c in text)
        return raw.decode("utf-8", errors="replace").replace("</w>", " ")
