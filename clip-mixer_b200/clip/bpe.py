"""Byte-pair encoding of raw strings for ``clip.tokenize`` (reference: ``SimpleTokenizer``,
training/clip/simple_tokenizer.py:62-132; host-side string processing, not on the GPU hot path).

The algorithm is the byte-level BPE the reference inherits from OpenAI CLIP, restated:
  * text clean-up: HTML-unescape twice, strip, collapse whitespace, lower-case (simple_tokenizer.py:50-59,
    :121; ``ftfy.fix_text`` is applied first only when ftfy is importable - it is absent in this image);
  * split with the CLIP pattern (contractions, letter runs, single digits, punctuation runs);
  * every UTF-8 byte maps to a printable code point; the last symbol of a word carries ``</w>``;
  * repeatedly merge the adjacent pair with the lowest rank in the merge table until none is ranked.
Vocabulary ids: 256 byte symbols, the same 256 with ``</w>``, the 48894 merges, then ``<|startoftext|>`` = 49406
and ``<|endoftext|>`` = 49407.

The merge table is a DATA asset of the reference (training/clip/bpe_simple_vocab_16e6.txt.gz).  It is looked up in
``$CLIP_BPE_VOCAB``, then ``baseline/_ref/`` (where ``__graft_entry__.build()`` copies it when /root/reference
exists; git-ignored).  Without the file ``encode`` raises: token ids can always be passed to ``tokenize`` directly.
Pinned against the reference's own tokenizer by tests/golden/bpe.json (oracle/make_golden_bpe.py)."""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import Dict, List, Tuple

from .._lib import MixerClipError

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_SEARCH = (os.path.join(_ROOT, "baseline", "_ref", "bpe_simple_vocab_16e6.txt.gz"),)
_N_MERGES = 49152 - 256 - 2


def vocab_path() -> str:
    cand = [os.environ.get("CLIP_BPE_VOCAB", "")] + list(_SEARCH)
    for p in cand:
        if p and os.path.isfile(p):
            return p
    raise MixerClipError(
        "tokenize() received a raw string but the BPE merge table was not found (set CLIP_BPE_VOCAB to the reference's "
        "bpe_simple_vocab_16e6.txt.gz, or pass token id sequences instead)")


@lru_cache()
def _byte_symbols() -> Dict[int, str]:
    """Printable stand-ins for the 256 byte values: the 188 printable Latin-1 code points keep their own character,
    the remaining 68 are assigned 256, 257, ... in byte order."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


class Tokenizer:
    def __init__(self, path: str = None):
        import regex
        path = path or vocab_path()
        with gzip.open(path, "rt", encoding="utf-8") as f:
            lines = f.read().split("\n")
        merges: List[Tuple[str, str]] = [tuple(l.split()) for l in lines[1:_N_MERGES + 1]]
        sym = _byte_symbols()
        # id order of the vocabulary: byte symbols sorted as the reference sorts them (the 188 kept bytes in byte
        # order, then the 68 re-mapped ones), their </w> forms, the merges, the two specials
        keep = [b for b in range(256) if ord(sym[b]) < 256]
        order = [sym[b] for b in keep] + [sym[b] for b in range(256) if ord(sym[b]) >= 256]
        vocab = order + [s + "</w>" for s in order] + ["".join(m) for m in merges] + ["<|startoftext|>", "<|endoftext|>"]
        self.encoder = {s: i for i, s in enumerate(vocab)}
        self.decoder = {i: s for s, i in self.encoder.items()}
        self.rank = {m: i for i, m in enumerate(merges)}
        self.sym = sym
        self.unsym = {v: k for k, v in sym.items()}
        self.cache = {"<|startoftext|>": ["<|startoftext|>"], "<|endoftext|>": ["<|endoftext|>"]}
        self.pat = regex.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                                 regex.IGNORECASE)

    def _merge_word(self, token: str) -> List[str]:
        hit = self.cache.get(token)
        if hit is not None:
            return hit
        parts = list(token[:-1]) + [token[-1] + "</w>"]
        while len(parts) > 1:
            best, best_rank = -1, None
            for i in range(len(parts) - 1):
                r = self.rank.get((parts[i], parts[i + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = i, r
            if best_rank is None:
                break
            a, b = parts[best], parts[best + 1]
            out, i = [], 0
            while i < len(parts):                     # merge EVERY occurrence of the winning pair, left to right
                if i + 1 < len(parts) and parts[i] == a and parts[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(parts[i])
                    i += 1
            parts = out
        self.cache[token] = parts
        return parts

    @staticmethod
    def clean(text: str) -> str:
        try:
            import ftfy
            text = ftfy.fix_text(text)
        except ImportError:
            pass
        text = html.unescape(html.unescape(text)).strip()
        return " ".join(text.split()).strip().lower()

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for word in self.pat.findall(self.clean(text)):
            mapped = "".join(self.sym[b] for b in word.encode("utf-8"))
            ids.extend(self.encoder[p] for p in self._merge_word(mapped))
        return ids

    def decode(self, ids) -> str:
        text = "".join(self.decoder[int(i)] for i in ids).replace("</w>", "\0")
        data = bytearray(32 if c == "\0" else self.unsym[c] for c in text if c == "\0" or c in self.unsym)
        return data.decode("utf-8", errors="replace")


_tok = None


def encode(text: str) -> List[int]:
    global _tok
    if _tok is None:
        _tok = Tokenizer()
    return _tok.encode(text)
