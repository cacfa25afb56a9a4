"""Byte-pair-encoding tokenizer compatible with CLIP's `bpe_simple_vocab_16e6.txt.gz`
(interface of TPT/clip/simple_tokenizer.py: `SimpleTokenizer().encode(text) -> List[int]`, `.decode`, `.encoder`).

The merge table is OpenAI's data file, not code, and is kept out of the git history (like the checkpoints).  It is
looked up in: an explicit `bpe_path`, $RLCF_BPE_VOCAB, this directory (where `__graft_entry__.build()` /
`baseline/make_ref.py` place the copy that ships next to the reference, TPT/clip/bpe_simple_vocab_16e6.txt.gz),
~/.cache/clip/.  Without it the tokenizer REFUSES to work (RuntimeError): token ids that do not match OpenAI's would
silently produce wrong class features with a real checkpoint.  The only exception is the explicit
`allow_byte_fallback=True` (set by clip.load("synthetic:...") for offline synthetic-weight runs, or
RLCF_BPE_FALLBACK=1): a byte-level vocabulary without merges whose sequences are still well-formed
(SOT ... EOT, ids < 49408).
"""
from __future__ import annotations

import gzip
import html
import os
import warnings
from functools import lru_cache

import regex as re

VOCAB_NAME = "bpe_simple_vocab_16e6.txt.gz"


def _find_vocab():
    cands = [os.environ.get("RLCF_BPE_VOCAB"), os.path.join(os.path.dirname(os.path.abspath(__file__)), VOCAB_NAME),
             os.path.expanduser(os.path.join("~/.cache/clip", VOCAB_NAME))]
    for c in cands:
        if c and os.path.isfile(c):
            return c
    return None


@lru_cache()
def bytes_to_unicode():
    """Reversible byte -> printable unicode character table used by GPT-2 style BPE vocabularies."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    chars = keep[:]
    extra = 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + extra)
            extra += 1
    return dict(zip(keep, (chr(c) for c in chars)))


def _clean(text: str) -> str:
    try:
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:  # ftfy only repairs mojibake; class names are plain ASCII
        pass
    text = html.unescape(html.unescape(text))
    return re.sub(r"\s+", " ", text.strip()).strip()


class SimpleTokenizer:
    def __init__(self, bpe_path: str | None = None, allow_byte_fallback: bool = False):
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        bpe_path = bpe_path or _find_vocab()
        if bpe_path is None and not (allow_byte_fallback or os.environ.get("RLCF_BPE_FALLBACK") == "1"):
            raise RuntimeError(
                f"{VOCAB_NAME} not found: set RLCF_BPE_VOCAB, put the file next to {os.path.abspath(__file__)} or in "
                "~/.cache/clip/ (`python baseline/make_ref.py` copies it from the reference).  Tokenising without "
                "OpenAI's merge table would give token ids that no CLIP checkpoint understands.")
        self.byte_fallback = bpe_path is None
        base = list(self.byte_encoder.values())
        vocab = base + [v + "</w>" for v in base]
        merges = []
        if bpe_path is not None:
            with gzip.open(bpe_path) as f:
                lines = f.read().decode("utf-8").split("\n")
            merges = [tuple(m.split()) for m in lines[1:49152 - 256 - 2 + 1]]
        else:
            warnings.warn(f"{VOCAB_NAME} not found (set RLCF_BPE_VOCAB): using a byte-level fallback vocabulary; token "
                          "ids will not match OpenAI CLIP's")
        vocab += ["".join(m) for m in merges]
        vocab += ["<|startoftext|>", "<|endoftext|>"]
        self.encoder = dict(zip(vocab, range(len(vocab))))
        if bpe_path is None:  # keep the special tokens at CLIP's ids so EOT stays the arg-max of every sequence
            self.encoder["<|startoftext|>"], self.encoder["<|endoftext|>"] = 49406, 49407
        self.decoder = {v: k for k, v in self.encoder.items()}
        self.ranks = dict(zip(merges, range(len(merges))))
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}
        self.pat = re.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                              re.IGNORECASE)

    def _bpe(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        word = list(token[:-1]) + [token[-1] + "</w>"]
        while len(word) > 1:
            pairs = [(self.ranks.get((a, b), float("inf")), i) for i, (a, b) in enumerate(zip(word, word[1:]))]
            rank, _ = min(pairs)
            if rank == float("inf"):
                break
            first, second = next((a, b) for (a, b) in zip(word, word[1:]) if self.ranks.get((a, b)) == rank)
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

    def encode(self, text: str):
        ids = []
        for token in re.findall(self.pat, _clean(text).lower()):
            token = "".join(self.byte_encoder[b] for b in token.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self._bpe(token).split(" "))
        return ids

    def decode(self, tokens):
        text = "".join(self.decoder[t] for t in tokens)
        return bytearray(self.byte_decoder[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")
