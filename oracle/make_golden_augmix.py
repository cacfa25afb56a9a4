"""Runs the REFERENCE's AugMixAugmenter (/root/reference/TPT/data/datautils.py) on seeded synthetic images and stores
SHA-256 digests + a few probe values of the returned views in tests/golden/augmix_ref.json (a few hundred bytes:
bit-exact pin for oracle/augmix_oracle.py).  Run here (the reference is not on the GPU box):
    python oracle/make_golden_augmix.py
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/TPT")
# data/datautils.py imports dataset builders the view generator does not use; stub the ones that need extra packages
for name in ("data.hoi_dataset",):
    m = types.ModuleType(name)
    m.BongardDataset = object
    sys.modules[name] = m

from oracle import augmix_oracle as A   # noqa: E402

CASES = [dict(name="in_a_like", h=375, w=500, seed=3, n_views=7, augmix=False),
         dict(name="tall", h=640, w=427, seed=4, n_views=5, augmix=False),
         dict(name="small_upscale", h=150, w=200, seed=5, n_views=5, augmix=False),
         dict(name="flowers_like", h=500, w=667, seed=6, n_views=9, augmix=True),
         dict(name="square_augmix", h=256, w=256, seed=7, n_views=12, augmix=True),
         dict(name="hard_aug_plain", h=375, w=500, seed=8, n_views=15, augmix=False, hard_aug=True),
         dict(name="hard_aug_augmix", h=333, w=250, seed=9, n_views=15, augmix=True, hard_aug=True)]


def digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def main():
    from data.datautils import AugMixAugmenter
    out = {}
    for c in CASES:
        img = A.synthetic_image(c["h"], c["w"], c["seed"])
        base, pre, _ = A.make_transforms()
        aug = AugMixAugmenter(base, pre, n_views=c["n_views"], augmix=c["augmix"], hard_aug=c.get("hard_aug", False))
        torch.manual_seed(c["seed"]); np.random.seed(c["seed"])
        views = torch.stack(aug(img))
        out[c["name"]] = dict(c, shape=list(views.shape), sha256=digest(views),
                              per_view=[digest(v)[:16] for v in views],
                              probes=[float(views[i % views.shape[0], i % 3, (37 * i) % 224, (91 * i) % 224]) for i in range(8)])
        print(c["name"], out[c["name"]]["sha256"][:16])
    with open(os.path.join(ROOT, "tests", "golden", "augmix_ref.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
