"""Byte-pair encoding of raw strings for ``clip.tokenize`` (host-side string processing, NOT on the
GPU hot path; SURVEY 2.1 row 7 marks it out of scope).  The 16e6-merge vocabulary file is a data
asset of the reference (training/clip/bpe_simple_vocab_16e6.txt.gz) and is not shipped here."""
import os

from .._lib import MixerClipError


def encode(text: str):
    path = os.environ.get("CLIP_BPE_VOCAB", "")
    raise MixerClipError(
        "tokenize() received a raw string but BPE encoding is not available in this build "
        f"(CLIP_BPE_VOCAB={path!r}); pass token id sequences instead")
