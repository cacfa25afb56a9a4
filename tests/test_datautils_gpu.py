"""On-device view generation (rlcf_b200/datautils.py + csrc/augment_kernels.cu through the C ABI) against the PIL
oracle: integer / byte work, so the bar is bit-exact -- both the uint8 crops and the final fp32 views."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import augmix_oracle as A

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augmix_ref.json")
with open(GOLDEN) as f:
    CASES = json.load(f)


@pytest.mark.parametrize("name", sorted(CASES))
def test_views_are_bit_exact_with_the_reference_pipeline(name):
    from rlcf_b200 import datautils as D
    c = CASES[name]
    img = A.synthetic_image(c["h"], c["w"], c["seed"])
    torch.manual_seed(c["seed"]); np.random.seed(c["seed"])
    hard = bool(c.get("hard_aug", False))
    ref = A.augmix_views(img, c["n_views"], bool(c["augmix"]), hard_aug=hard)
    torch.manual_seed(c["seed"]); np.random.seed(c["seed"])
    aug = D.AugMixAugmenter(None, None, n_views=c["n_views"], augmix=bool(c["augmix"]), hard_aug=hard, device=DEV)
    views = aug(img)
    assert isinstance(views, list) and len(views) == c["n_views"] + 1 and views[0].is_cuda
    got = torch.stack(views).cpu()
    bad = [i for i in range(got.shape[0]) if not torch.equal(got[i], ref[i])]
    assert not bad, f"views {bad} differ from the reference pipeline; max |diff| {(got - ref).abs().max():.3e}"
    import hashlib
    assert hashlib.sha256(got.contiguous().numpy().tobytes()).hexdigest() == c["sha256"]   # the reference's own digest


def test_sixty_four_views_feed_the_engine_shape():
    """The shape the hot path consumes: 1 + 63 views of one ImageNet-sized image, one small upload."""
    from rlcf_b200 import datautils as D
    img = A.synthetic_image(375, 500, 21)
    torch.manual_seed(1); np.random.seed(1)
    ref = A.augmix_views(img, 63, False)
    torch.manual_seed(1); np.random.seed(1)
    v = D.AugMixAugmenter(n_views=63, augmix=False, device=DEV).views(img)
    assert tuple(v.shape) == (64, 3, 224, 224) and v.dtype == torch.float32
    assert torch.equal(v.cpu(), ref)


def test_bad_inputs_are_rejected():
    from rlcf_b200 import datautils as D
    from rlcf_b200._lib import RlcfError
    with pytest.raises(RlcfError):
        D.AugMixAugmenter(n_views=3, device=DEV).views(torch.zeros(10, 10, 3))   # not uint8


@pytest.mark.parametrize("hard_aug", [0, 1])
def test_driver_on_an_image_folder_with_gpu_views(tmp_path, hard_aug):
    """tune_cls_rl.main_worker on a real ImageFolder (PNG files decoded by PIL, views generated on the GPU, batched
    LayerNorm-tuning RLCF) with seeded random-init ViT-B/32 weights: the whole reference flow minus the checkpoints."""
    from rlcf_b200 import params, tune_cls_rl
    root = tmp_path / "data" / "imagenet-a"
    names = ["n01_goldfish", "n02_tabby_cat", "n03_fire_truck", "n04_volcano", "n05_accordion", "n06_snail"]
    for ci, cname in enumerate(names):     # top-5 accuracy needs at least 5 classes (tools.accuracy, as the reference)
        (root / cname).mkdir(parents=True)
        A.synthetic_image(120 + 10 * ci, 160 - 7 * ci, 100 + ci).save(root / cname / "img0.png")
    args = params.build_parser().parse_args([
        str(tmp_path / "data"), "--test_sets", "A", "-a", "ViT-B/32", "--reward_arch", "ViT-B/32", "--tpt",
        "--tune_norm", "1", "--batch_size", "8", "--selection_p", "0.5", "--tta_steps", "1", "--sample_k", "2",
        "--synthetic_weights", "--images_per_step", "2", "--workers", "0", "--output", str(tmp_path / "out"),
        "--hard_aug", str(hard_aug)])
    os.makedirs(args.output, exist_ok=True)
    res = tune_cls_rl.main_worker(0, args)
    top1, top5 = res["A"]
    assert 0.0 <= top1 <= top5 <= 100.0
    assert args.n_images == 6 and args.n_classes == 6


def test_device_tap_tables_equal_the_host_restatement_of_pillow():
    """rlcf_resample_taps (device, double precision) against datautils.resample_taps_batch (numpy restatement of
    Pillow's precompute_coeffs, itself pinned to PIL in tests/test_datautils_host_cpu.py): bit-identical tables."""
    from rlcf_b200 import datautils as D, ops
    for (w, h, seed) in ((500, 375, 1), (427, 640, 2), (200, 150, 3), (3000, 2000, 4), (224, 224, 5)):
        torch.manual_seed(seed); np.random.seed(seed)
        host = D.sample_plan(w, h, 15, False, host_taps=True)
        torch.manual_seed(seed); np.random.seed(seed)
        dev = D.sample_plan(w, h, 15, False, host_taps=False)
        hdr = torch.from_numpy(dev.hdr).to(DEV)
        hb, hk, vb, vk = ops.resample_taps(torch.from_numpy(dev.geom).to(DEV), 224, dev.ks_h, dev.ks_v, hdr)
        assert np.array_equal(hdr.cpu().numpy(), host.hdr)
        assert np.array_equal(hb.cpu().numpy(), host.hb) and np.array_equal(vb.cpu().numpy(), host.vb)
        for got, want in ((hk.cpu().numpy(), host.hk), (vk.cpu().numpy(), host.vk)):
            k = min(got.shape[2], want.shape[2])
            assert np.array_equal(got[:, :, :k], want[:, :, :k])
            assert not got[:, :, k:].any() and not want[:, :, k:].any()
