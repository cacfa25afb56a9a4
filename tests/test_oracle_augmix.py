"""Pins oracle/augmix_oracle.py (the CPU restatement of the reference's view generation) to digests of what the
reference's own AugMixAugmenter returned (tests/golden/augmix_ref.json, oracle/make_golden_augmix.py).  No GPU."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import augmix_oracle as A

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augmix_ref.json")


def cases():
    with open(GOLDEN) as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(cases()))
def test_view_generation_oracle_is_bit_exact_with_reference(name):
    c = cases()[name]
    img = A.synthetic_image(c["h"], c["w"], c["seed"])
    torch.manual_seed(c["seed"]); np.random.seed(c["seed"])
    views = A.augmix_views(img, c["n_views"], bool(c["augmix"]), hard_aug=bool(c.get("hard_aug", False)))
    assert list(views.shape) == c["shape"]
    per_view = [hashlib.sha256(v.contiguous().numpy().tobytes()).hexdigest()[:16] for v in views]
    assert per_view == c["per_view"]
    assert hashlib.sha256(views.contiguous().numpy().tobytes()).hexdigest() == c["sha256"]
