"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's view generation (TPT/data/datautils.py:76-128,
TPT/data/augmix_ops.py, TPT/tune_cls_rl.py:102-110) on PIL + torchvision, the libraries the reference itself calls.
Pinned to the reference: tests/golden/augmix_*.json holds SHA-256 digests of what /root/reference's AugMixAugmenter
returned for seeded synthetic images (oracle/make_golden_augmix.py); tests/test_oracle_augmix.py holds this file to
them bit for bit.  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torchvision.transforms as T
from PIL import Image, ImageOps
from torchvision.transforms import InterpolationMode

MEAN = [0.48145466, 0.4578275, 0.40821073]      # tune_cls_rl.py:92-93
STD = [0.26862954, 0.26130258, 0.27577711]
IMAGE_SIZE = 224                                # augmix_ops.py:6


def synthetic_image(h: int, w: int, seed: int) -> Image.Image:
    """A seeded RGB test image with smooth structure + noise (so that resampling, histograms and LUTs all matter)."""
    g = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        fx, fy, ph = g.uniform(0.01, 0.08), g.uniform(0.01, 0.08), g.uniform(0, 6.28)
        img[:, :, c] = 120 + 90 * np.sin(fx * xx + fy * yy + ph) + g.normal(0, 18, size=(h, w))
    return Image.fromarray(np.clip(img, 0, 255).astype(np.uint8), "RGB")


# ---- augmix_ops.py:9-107 (the nine operations of `augmentations`, augmix_ops.py:141-144) --------------------------
def _int_parameter(level, maxval):
    return int(level * maxval / 10)


def _float_parameter(level, maxval):
    return float(level) * maxval / 10.


def _sample_level(n):
    return np.random.uniform(low=0.1, high=n)


def _autocontrast(img, _):
    return ImageOps.autocontrast(img)


def _equalize(img, _):
    return ImageOps.equalize(img)


def _posterize(img, level):
    return ImageOps.posterize(img, 4 - _int_parameter(_sample_level(level), 4))


def _rotate(img, level):
    degrees = _int_parameter(_sample_level(level), 30)
    if np.random.uniform() > 0.5:
        degrees = -degrees
    return img.rotate(degrees, resample=Image.BILINEAR)


def _solarize(img, level):
    return ImageOps.solarize(img, 256 - _int_parameter(_sample_level(level), 256))


def _affine(img, coeffs):
    return img.transform((IMAGE_SIZE, IMAGE_SIZE), Image.AFFINE, coeffs, resample=Image.BILINEAR)


def _shear_x(img, level):
    level = _float_parameter(_sample_level(level), 0.3)
    if np.random.uniform() > 0.5:
        level = -level
    return _affine(img, (1, level, 0, 0, 1, 0))


def _shear_y(img, level):
    level = _float_parameter(_sample_level(level), 0.3)
    if np.random.uniform() > 0.5:
        level = -level
    return _affine(img, (1, 0, 0, level, 1, 0))


def _translate_x(img, level):
    level = _int_parameter(_sample_level(level), IMAGE_SIZE / 3)
    if np.random.random() > 0.5:
        level = -level
    return _affine(img, (1, 0, level, 0, 1, 0))


def _translate_y(img, level):
    level = _int_parameter(_sample_level(level), IMAGE_SIZE / 3)
    if np.random.random() > 0.5:
        level = -level
    return _affine(img, (1, 0, 0, 0, 1, level))


AUGMENTATIONS = [_autocontrast, _equalize, _posterize, _rotate, _solarize, _shear_x, _shear_y, _translate_x, _translate_y]


def make_transforms(resolution: int = 224, hard_aug: bool = False, crop_min: float = 0.2):
    """tune_cls_rl.py:102-108 and get_preaugment (datautils.py:76-92)."""
    base = T.Compose([T.Resize(resolution, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(resolution)])
    pre = T.Compose([T.ToTensor(), T.Normalize(mean=MEAN, std=STD)])
    if hard_aug:    # datautils.py:77-87
        preaug = T.Compose([T.RandomResizedCrop(resolution, scale=(crop_min, 1.)),
                            T.RandomApply([T.ColorJitter(0.4, 0.4, 0.2, 0.1)], p=0.5), T.RandomGrayscale(p=0.2),
                            T.RandomApply([T.GaussianBlur(3, sigma=(0.1, 2.0))], p=0.1), T.RandomHorizontalFlip()])
    else:
        preaug = T.Compose([T.RandomResizedCrop(224), T.RandomHorizontalFlip()])
    return base, pre, preaug


def augmix_view(image, preaugment, preprocess, aug_list, severity=1):
    """datautils.augmix (datautils.py:95-111)."""
    x_orig = preaugment(image)
    x_processed = preprocess(x_orig)
    if len(aug_list) == 0:
        return x_processed
    w = np.float32(np.random.dirichlet([1.0, 1.0, 1.0]))
    m = np.float32(np.random.beta(1.0, 1.0))
    mix = torch.zeros_like(x_processed)
    for i in range(3):
        x_aug = x_orig.copy()
        for _ in range(np.random.randint(1, 4)):
            x_aug = AUGMENTATIONS[np.random.choice(len(aug_list))](x_aug, severity)
        mix += w[i] * preprocess(x_aug)
    return m * x_processed + (1 - m) * mix


def augmix_views(image, n_views: int, augmix: bool, severity: int = 1, hard_aug: bool = False) -> torch.Tensor:
    """AugMixAugmenter.__call__ (datautils.py:114-128): [image] + n_views augmented views, stacked [n_views+1,3,224,224].
    Randomness comes from the global torch and numpy generators, exactly as in the reference."""
    base, pre, preaug = make_transforms(hard_aug=hard_aug)
    aug_list = AUGMENTATIONS if augmix else []
    out = [pre(base(image))]
    out += [augmix_view(image, preaug, pre, aug_list, severity) for _ in range(n_views)]
    return torch.stack(out)
