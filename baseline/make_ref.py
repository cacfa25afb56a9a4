"""Recipe for the reference arm: copies the UNMODIFIED reference sources of the hot path (/root/reference/TPT, 1.7 MB,
pure Python + the BPE vocabulary) into the git-ignored directory baseline/_ref/TPT so that they travel to the GPU box
with the gpurun snapshot, exactly like the built librlcf_b200.so does.  Nothing under baseline/_ref/ is product code or
is ever committed; rlcf_b200/ never imports it (only bench.py's baseline legs and baseline/ref_harness.py do).

    python baseline/make_ref.py            # also called by __graft_entry__.build() when /root/reference exists

The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` is not
possible; a verbatim tree copy is the equivalent.  A manifest with the SHA-256 of every copied
file is written next to the copy so that a reader can check that the files are unmodified.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/TPT"
DST = os.path.join(HERE, "_ref", "TPT")


def make_ref(src: str = SRC, dst: str = DST) -> bool:
    """Returns True when baseline/_ref/TPT is in place (copied now or earlier), False when the reference is absent
    (e.g. on the GPU box, where the prebuilt copy is used)."""
    if not os.path.isdir(src):
        return os.path.isdir(dst)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for root, _, files in os.walk(dst):
        for f in sorted(files):
            p = os.path.join(root, f)
            with open(p, "rb") as fh:
                manifest[os.path.relpath(p, dst)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(os.path.dirname(dst), "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    # OpenAI's BPE merge table (a data file, like the checkpoints): placed next to rlcf_b200's tokenizer, git-ignored
    vocab = os.path.join(src, "clip", "bpe_simple_vocab_16e6.txt.gz")
    if os.path.isfile(vocab):
        shutil.copyfile(vocab, os.path.join(os.path.dirname(HERE), "rlcf_b200", "clip", "bpe_simple_vocab_16e6.txt.gz"))
    return True


if __name__ == "__main__":
    ok = make_ref()
    print("baseline/_ref/TPT", "ready" if ok else "NOT available (no /root/reference and no earlier copy)")
    sys.exit(0 if ok else 1)
