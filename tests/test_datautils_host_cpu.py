"""Host side of the on-device view generation (rlcf_b200/datautils.py): the random plan + Pillow tap tables, executed
by a numpy re-statement of the kernels (tests/plan_numpy.py), must reproduce the PIL oracle bit for bit.  No GPU."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import augmix_oracle as A
import plan_numpy as PN

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augmix_ref.json")
with open(GOLDEN) as f:
    CASES = json.load(f)


@pytest.mark.parametrize("name", sorted(CASES))
def test_plan_reproduces_the_reference_views(name):
    from rlcf_b200 import datautils as D
    c = CASES[name]
    img = A.synthetic_image(c["h"], c["w"], c["seed"])
    torch.manual_seed(c["seed"]); np.random.seed(c["seed"])
    hard = bool(c.get("hard_aug", False))
    ref = A.augmix_views(img, c["n_views"], bool(c["augmix"]), hard_aug=hard).numpy()
    torch.manual_seed(c["seed"]); np.random.seed(c["seed"])
    if hard:
        plan = D.sample_plan_hard(img, c["n_views"], bool(c["augmix"]))
    else:
        plan = D.sample_plan(c["w"], c["h"], c["n_views"], bool(c["augmix"]))
    got = PN.execute(np.asarray(img), plan)
    bad = np.argwhere((got != ref).reshape(got.shape[0], -1).any(axis=1)).ravel().tolist()
    assert not bad, f"views {bad} differ; max |diff| {np.abs(got - ref).max():.3e}"


def test_taps_are_pillows_for_plain_resizes():
    """resample_taps against PIL itself on a whole-image resize (both filters, down- and up-scaling)."""
    from PIL import Image
    from rlcf_b200 import datautils as D
    img = A.synthetic_image(97, 131, 11)
    src = np.asarray(img)
    for (ow, oh), filt, pil_f in (((50, 40), "bilinear", Image.BILINEAR), ((224, 160), "bicubic", Image.BICUBIC),
                                  ((131, 30), "bilinear", Image.BILINEAR), ((33, 200), "bicubic", Image.BICUBIC)):
        ref = np.asarray(img.resize((ow, oh), pil_f))
        hb, hk = D.resample_taps(131, ow, filt)
        vb, vk = D.resample_taps(97, oh, filt)
        tmp = np.zeros((97, ow, 3), dtype=np.uint8)
        for x in range(ow):
            k = hk[x, :hb[x, 1]].astype(np.int64)
            s = (1 << 21) + (src[:, hb[x, 0]:hb[x, 0] + hb[x, 1], :].astype(np.int64) * k[None, :, None]).sum(1)
            tmp[:, x] = PN.clip8(s)
        out = np.zeros((oh, ow, 3), dtype=np.uint8)
        for y in range(oh):
            k = vk[y, :vb[y, 1]].astype(np.int64)
            s = (1 << 21) + (tmp[vb[y, 0]:vb[y, 0] + vb[y, 1]].astype(np.int64) * k[:, None, None]).sum(0)
            out[y] = PN.clip8(s)
        assert np.array_equal(out, ref), (ow, oh, filt)
