#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/bpe.json by running the REAL reference tokenizer
(/root/reference/training/clip/simple_tokenizer.py + clip.tokenize's padding rule, clip.py:198-238) on a fixed list of
strings.  ``ftfy`` is absent from this image, so the reference module is imported with an identity ``fix_text`` stub
(all test strings are already well-formed Unicode, for which ftfy.fix_text is the identity)."""
import importlib.util
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/training/clip/simple_tokenizer.py"

STRINGS = [
    "a photo of a cat", "A diagram.", "a bad photo of a tench, Tinca tinca", "itap of the goldfish.",
    "Hello,   World!! it's 2023 &amp; I'm fine", "naïve café — ünïcödé 😀 test", "don't they've we'll he'd I'M",
    "12345 67.89 $9.99 (50% off)", "supercalifragilisticexpialidocious antidisestablishmentarianism",
    "  leading and trailing   ", "", "日本語のテキスト and русский текст", "a" * 300,
]


def main():
    ft = types.ModuleType("ftfy")
    ft.fix_text = lambda t: t
    sys.modules["ftfy"] = ft
    spec = importlib.util.spec_from_file_location("ref_simple_tokenizer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tok = mod.SimpleTokenizer()
    sot, eot = tok.encoder["<|startoftext|>"], tok.encoder["<|endoftext|>"]
    rows = []
    for s in STRINGS:
        ids = tok.encode(s)
        full = [sot] + ids + [eot]
        if len(full) > 77:                      # clip.py:230-234 with truncate=True
            full = full[:77]
            full[-1] = eot
        rows.append({"text": s, "ids": ids, "tokenized_truncate": full + [0] * (77 - len(full))})
    out = os.path.join(ROOT, "tests", "golden", "bpe.json")
    json.dump({"source": "reference SimpleTokenizer (simple_tokenizer.py:62-132), ftfy.fix_text = identity", "rows": rows},
              open(out, "w"), ensure_ascii=True, indent=0)
    print(out, len(rows))


if __name__ == "__main__":
    main()
