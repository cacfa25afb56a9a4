"""View generation with the reference's surface (TPT/data/datautils.py:76-128) executed on the GPU.

The reference builds, per test image, the un-augmented view (Resize(224, BICUBIC) + CenterCrop) and `n_views`
augmented ones (RandomResizedCrop(224) + RandomHorizontalFlip, then AugMix chains for the fine-grained datasets,
tune_cls_rl.py:102-110) with PIL on the host.  Here the host only draws the random decisions -- from the same global
torch / numpy generators, in the same order, so a seeded run produces the reference's views -- and precomputes what
Pillow precomputes per call (resampling taps, affine coefficients); the pixels are produced by
librlcf_b200.so (csrc/augment_kernels.cu) from ONE upload of the decoded uint8 image.

    aug = AugMixAugmenter(None, None, n_views=63, augmix=False)     # same constructor as the reference
    views = aug(pil_image)            # list of 64 device tensors [3,224,224] (views of one [64,3,224,224] tensor)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import ops
from ._lib import RlcfError

MEAN = (0.48145466, 0.4578275, 0.40821073)      # tune_cls_rl.py:92-93
STD = (0.26862954, 0.26130258, 0.27577711)
OUT = 224
PRECISION_BITS = 32 - 8 - 2                      # Pillow, libImaging/Resample.c

OP_AUTOCONTRAST, OP_EQUALIZE, OP_POSTERIZE, OP_SOLARIZE, OP_AFFINE = range(5)
# order of augmix_ops.augmentations (augmix_ops.py:141-144)
AUG_NAMES = ("autocontrast", "equalize", "posterize", "rotate", "solarize", "shear_x", "shear_y", "translate_x",
             "translate_y")


# ------------------------------------------------------------------------------------------------ Pillow's resampling taps
def _bilinear(x):
    x = np.abs(x)
    return np.where(x < 1.0, 1.0 - x, 0.0)


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1,
                    np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


FILTERS = {"bilinear": (_bilinear, 1.0), "bicubic": (_bicubic, 2.0)}


def resample_taps_batch(in_sizes, out_size: int, filt: str, out_lo: int = 0, out_n: int | None = None):
    """precompute_coeffs + normalize_coeffs_8bpc (libImaging/Resample.c) for resizing in_sizes[v] samples to `out_size`
    (all views at once), restricted to outputs [out_lo, out_lo + out_n).  Returns (bounds int32 [V,out_n,2] = (first tap,
    count), taps int32 [V,out_n,ks_max]).  Same double-precision operation order as the C source; the taps of one
    output are accumulated sequentially."""
    fn, support0 = FILTERS[filt]
    out_n = out_size if out_n is None else out_n
    in_sizes = np.asarray(in_sizes, dtype=np.int64).reshape(-1)
    scale = ((in_sizes.astype(np.float32) - np.float32(0)).astype(np.float64) / out_size)[:, None]     # [V,1]
    filterscale = np.maximum(scale, 1.0)
    support = support0 * filterscale
    ksize = int(np.ceil(support).max()) * 2 + 1
    xx = np.arange(out_lo, out_lo + out_n, dtype=np.float64)[None, :]
    center = 0.0 + (xx + 0.5) * scale                                       # [V,out_n]
    ss = 1.0 / filterscale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)         # (int) truncates toward zero, then clamp
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_sizes[:, None]) - xmin
    k = np.zeros(center.shape + (ksize,), dtype=np.float64)
    ww = np.zeros_like(center)
    for x in range(ksize):
        w = np.where(x < xmax, fn((x + xmin - center + 0.5) * ss), 0.0)
        k[..., x] = w
        ww = ww + w
    nz = ww != 0.0
    k[nz] = k[nz] / ww[nz][:, None]
    fixed = np.where(k < 0, -0.5 + k * (1 << PRECISION_BITS), 0.5 + k * (1 << PRECISION_BITS)).astype(np.int64)
    bounds = np.stack([xmin, xmax], axis=-1).astype(np.int32)
    return bounds, fixed.astype(np.int32)


def resample_taps(in_size: int, out_size: int, filt: str, out_lo: int = 0, out_n: int | None = None):
    """Single-view form of resample_taps_batch: (bounds [out_n,2], taps [out_n,ksize])."""
    b, t = resample_taps_batch([in_size], out_size, filt, out_lo, out_n)
    return b[0], t[0]


# ------------------------------------------------------------------------------------------------ the random plan
@dataclass
class ViewPlan:
    """Everything the kernels need for the views of one image (host numpy arrays)."""
    hdr: np.ndarray      # int32 [V,8]   x0, y0, row_first, n_rows, flip
    hb: np.ndarray       # int32 [V,224,2]
    hk: np.ndarray       # int32 [V,224,ks_h]
    vb: np.ndarray       # int32 [V,224,2]
    vk: np.ndarray       # int32 [V,224,ks_v]
    vflag: np.ndarray    # int32 [V]     1 = AugMix blend, 0 = plain normalised view
    wts: np.ndarray      # float32 [V,4] w0, w1, w2, m
    omm: np.ndarray      # float32 [V]   float32(1 - m)
    n_ops: np.ndarray    # int32 [V,3]
    ops: np.ndarray      # int32 [V,3,3,2]
    mats: np.ndarray     # float64 [V,3,3,6]
    tmp_rows: int
    geom: np.ndarray | None = None   # int32 [V,8] in_w, in_h, res_w, res_h, lo_x, lo_y, filter: taps are then computed
    ks_h: int = 0                    # on the device (hb/hk/vb/vk are None) with these tap-count bounds
    ks_v: int = 0
    x_pre: np.ndarray | None = None  # uint8 [V-1,224,224,3]: hard_aug pre-augmented crops made on the host (views 1..)


def _window(vb):
    """Rows of the source the vertical pass touches (ImagingResample: ybox_first / ybox_last) and the row taps
    re-based to the first of them.  vb [..., out, 2]."""
    first = vb[..., 0, 0].copy()
    n_rows = vb[..., -1, 0] + vb[..., -1, 1] - first
    vb = vb.copy()
    vb[..., 0] -= first[..., None]
    return vb, first, n_rows


def _rotate_matrix(degrees: int):
    """Image.rotate(angle, BILINEAR) on a 224 x 224 image (PIL/Image.py): the AFFINE coefficients it hands to
    Image.transform, or None when the fast path returns a copy."""
    angle = degrees % 360.0
    if angle == 0:
        return None
    w = h = OUT
    cx, cy = w / 2, h / 2
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2]
    m[5] = m[3] * (-cx) + m[4] * (-cy) + m[5]
    m[2] += cx
    m[5] += cy
    return m


def _sample_op(name: str, severity):
    """Draws one operation's parameters exactly as augmix_ops.py:44-107 does and returns (type, int param, matrix)."""
    if name == "autocontrast":
        return OP_AUTOCONTRAST, 0, None
    if name == "equalize":
        return OP_EQUALIZE, 0, None
    level = np.random.uniform(low=0.1, high=severity)                    # sample_level
    if name == "posterize":
        return OP_POSTERIZE, 4 - int(level * 4 / 10), None
    if name == "solarize":
        return OP_SOLARIZE, 256 - int(level * 256 / 10), None
    if name == "rotate":
        degrees = int(level * 30 / 10)
        if np.random.uniform() > 0.5:
            degrees = -degrees
        m = _rotate_matrix(degrees)
        return (OP_AFFINE, 0, m) if m is not None else (-1, 0, None)
    if name in ("shear_x", "shear_y"):
        lv = float(level) * 0.3 / 10.
        if np.random.uniform() > 0.5:
            lv = -lv
        return OP_AFFINE, 0, ([1, lv, 0, 0, 1, 0] if name == "shear_x" else [1, 0, 0, lv, 1, 0])
    if name in ("translate_x", "translate_y"):
        lv = int(level * (OUT / 3) / 10)
        if np.random.random() > 0.5:
            lv = -lv
        return OP_AFFINE, 0, ([1, 0, lv, 0, 1, 0] if name == "translate_x" else [1, 0, 0, 0, 1, lv])
    raise KeyError(name)


def _sample_augmix(v, severity, vflag, wts, omm, n_ops, op_codes, mats):
    """The numpy draws of augmix() for view v (datautils.py:95-111), written into row v of the plan tables."""
    vflag[v] = 1
    ws = np.float32(np.random.dirichlet([1.0, 1.0, 1.0]))
    m = np.float32(np.random.beta(1.0, 1.0))
    wts[v, :3], wts[v, 3] = ws, m
    omm[v] = 1 - m
    for c in range(3):
        depth = np.random.randint(1, 4)
        k = 0
        for _ in range(depth):
            name = AUG_NAMES[np.random.choice(len(AUG_NAMES))]
            t, p, mat = _sample_op(name, severity)
            if t < 0:          # rotate by 0 degrees: Image.rotate returns a copy
                continue
            op_codes[v, c, k] = (t, p)
            if mat is not None:
                mats[v, c, k] = mat
            k += 1
        n_ops[v, c] = k


def hard_preaugment(resolution: int = OUT, crop_min: float = 0.2):
    """get_preaugment(hard_aug=True) (datautils.py:76-87): the BYOL recipe.  It is PIL work on the data-loader side
    (colour jitter in HSV, grayscale, Gaussian blur), so it stays torchvision on the host -- same calls, same draws
    from the global torch generator as the reference; the AugMix chains and the normalisation still run on the GPU."""
    import torchvision.transforms as T
    return T.Compose([T.RandomResizedCrop(resolution, scale=(crop_min, 1.)),
                      T.RandomApply([T.ColorJitter(0.4, 0.4, 0.2, 0.1)], p=0.5),
                      T.RandomGrayscale(p=0.2),
                      T.RandomApply([T.GaussianBlur(3, sigma=(0.1, 2.0))], p=0.1),
                      T.RandomHorizontalFlip()])


def sample_plan_hard(image, n_views: int, augmix: bool, severity: int = 1, host_taps: bool = True) -> ViewPlan:
    """sample_plan for AugMixAugmenter(hard_aug=True): `image` is the PIL image; views 1.. are pre-augmented on the
    host (hard_preaugment) into plan.x_pre, view 0 and everything after the pre-augmentation is device work."""
    from PIL import Image
    if not isinstance(image, Image.Image):
        image = Image.fromarray(_to_u8_hwc(image).numpy())
    image = image.convert("RGB")
    plan = sample_plan(image.size[0], image.size[1], 0, False, severity, host_taps)       # view 0: no random draws
    V = n_views + 1
    pre = hard_preaugment()
    x_pre = np.zeros((n_views, OUT, OUT, 3), dtype=np.uint8)
    vflag = np.zeros(V, dtype=np.int32)
    wts = np.zeros((V, 4), dtype=np.float32)
    omm = np.zeros(V, dtype=np.float32)
    n_ops = np.zeros((V, 3), dtype=np.int32)
    op_codes = np.zeros((V, 3, 3, 2), dtype=np.int32)
    mats = np.zeros((V, 3, 3, 6), dtype=np.float64)
    for v in range(1, V):
        x_pre[v - 1] = np.asarray(pre(image))
        if augmix:
            _sample_augmix(v, severity, vflag, wts, omm, n_ops, op_codes, mats)
    plan.vflag, plan.wts, plan.omm, plan.n_ops, plan.ops, plan.mats, plan.x_pre = vflag, wts, omm, n_ops, op_codes, mats, x_pre
    return plan


def _ksize(in_sizes, res_sizes, support0):
    """Largest tap count of a set of resizes: ceil(support * max(scale, 1)) * 2 + 1 (precompute_coeffs)."""
    scale = np.asarray(in_sizes, dtype=np.float64) / np.asarray(res_sizes, dtype=np.float64)
    return int(np.ceil(support0 * np.maximum(scale, 1.0)).max()) * 2 + 1


def sample_plan(img_w: int, img_h: int, n_views: int, augmix: bool, severity: int = 1, host_taps: bool = True) -> ViewPlan:
    """The random decisions of AugMixAugmenter.__call__ (datautils.py:123-126) for one image of img_w x img_h pixels,
    drawn from the global torch / numpy generators in the reference's order, plus the derived tables.
    host_taps=False leaves Pillow's tap tables to the device (rlcf_resample_taps) and only records the geometry."""
    from torchvision.transforms import RandomResizedCrop
    V = n_views + 1
    hdr = np.zeros((V, 8), dtype=np.int32)
    vflag = np.zeros(V, dtype=np.int32)
    wts = np.zeros((V, 4), dtype=np.float32)
    omm = np.zeros(V, dtype=np.float32)
    n_ops = np.zeros((V, 3), dtype=np.int32)
    op_codes = np.zeros((V, 3, 3, 2), dtype=np.int32)
    mats = np.zeros((V, 3, 3, 6), dtype=np.float64)
    # view 0: transforms.Resize(224, BICUBIC) + CenterCrop(224)            (tune_cls_rl.py:103-105): the window of
    # OUT x OUT outputs at (left, top) of the image resized to new_w x new_h
    if img_w <= img_h:
        new_w, new_h = OUT, int(OUT * img_h / img_w)
    else:
        new_w, new_h = int(OUT * img_w / img_h), OUT
    top, left = int(round((new_h - OUT) / 2.0)), int(round((new_w - OUT) / 2.0))
    hb0, hk0 = resample_taps(img_w, new_w, "bicubic", left, OUT)
    vb0, vk0 = resample_taps(img_h, new_h, "bicubic", top, OUT)
    vb0, rf0, nr0 = _window(vb0)
    hdr[0, :5] = (0, 0, rf0, nr0, 0)
    dummy = torch.empty(3, img_h, img_w, dtype=torch.uint8)
    crop_w, crop_h = np.zeros(n_views, dtype=np.int64), np.zeros(n_views, dtype=np.int64)
    for v in range(1, V):
        # preaugment: RandomResizedCrop(224) + RandomHorizontalFlip           (datautils.py:89-92)
        i, j, h, w = RandomResizedCrop.get_params(dummy, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0))
        flip = bool(torch.rand(1) < 0.5)
        hdr[v, :5] = (j, i, 0, 0, int(flip))
        crop_w[v - 1], crop_h[v - 1] = w, h
        if not augmix:
            continue
        _sample_augmix(v, severity, vflag, wts, omm, n_ops, op_codes, mats)
    if not host_taps:
        geom = np.zeros((V, 8), dtype=np.int32)
        geom[0, :7] = (img_w, img_h, new_w, new_h, left, top, 1)
        geom[1:, 0], geom[1:, 1], geom[1:, 2], geom[1:, 3] = crop_w, crop_h, OUT, OUT
        ks_h = max(_ksize([img_w], [new_w], 2.0), _ksize(crop_w, OUT, 1.0) if n_views else 0)
        ks_v = max(_ksize([img_h], [new_h], 2.0), _ksize(crop_h, OUT, 1.0) if n_views else 0)
        hdr[0, 2:4] = 0
        return ViewPlan(hdr=hdr, hb=None, hk=None, vb=None, vk=None, vflag=vflag, wts=wts, omm=omm, n_ops=n_ops,
                        ops=op_codes, mats=mats, tmp_rows=img_h, geom=geom, ks_h=ks_h, ks_v=ks_v)
    # taps of all random crops at once (the crop is resized as an image of its own: F.resized_crop)
    hb = np.zeros((V, OUT, 2), dtype=np.int32)
    vb = np.zeros((V, OUT, 2), dtype=np.int32)
    hb[0], vb[0] = hb0, vb0
    ks_h, ks_v = hk0.shape[1], vk0.shape[1]
    if n_views > 0:
        hbn, hkn = resample_taps_batch(crop_w, OUT, "bilinear")
        vbn, vkn = resample_taps_batch(crop_h, OUT, "bilinear")
        vbn, rfn, nrn = _window(vbn)
        hb[1:], vb[1:] = hbn, vbn
        hdr[1:, 2], hdr[1:, 3] = rfn, nrn
        ks_h, ks_v = max(ks_h, hkn.shape[2]), max(ks_v, vkn.shape[2])
    hk = np.zeros((V, OUT, ks_h), dtype=np.int32)
    vk = np.zeros((V, OUT, ks_v), dtype=np.int32)
    hk[0, :, :hk0.shape[1]], vk[0, :, :vk0.shape[1]] = hk0, vk0
    if n_views > 0:
        hk[1:, :, :hkn.shape[2]], vk[1:, :, :vkn.shape[2]] = hkn, vkn
    return ViewPlan(hdr=hdr, hb=hb, hk=hk, vb=vb, vk=vk, vflag=vflag, wts=wts, omm=omm, n_ops=n_ops, ops=op_codes,
                    mats=mats, tmp_rows=int(hdr[:, 3].max()))


# ------------------------------------------------------------------------------------------------ device execution
def _to_u8_hwc(x) -> torch.Tensor:
    """PIL image / numpy array / tensor -> contiguous uint8 [H,W,3] (host)."""
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(x)
    else:   # PIL
        t = torch.from_numpy(np.asarray(x.convert("RGB")).copy())
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise RlcfError(f"expected a uint8 [H,W,3] image, got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


def run_plan(image_u8: torch.Tensor, plan: ViewPlan, device, out: torch.Tensor | None = None) -> torch.Tensor:
    """Executes a ViewPlan on `device`: uint8 [H,W,3] image -> fp32 views [V,3,224,224] (written into `out` when given,
    e.g. this image's slice of an engine's input batch)."""
    dev = torch.device(device)
    src = image_u8.to(dev, non_blocking=True)
    Vr = plan.hdr.shape[0]                          # views the resampling kernels produce (1 with hard_aug)
    V = Vr if plan.x_pre is None else 1 + plan.x_pre.shape[0]

    def d(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
    tmp = torch.empty(Vr, plan.tmp_rows, OUT, 3, dtype=torch.uint8, device=dev)
    x_orig = torch.empty(V, OUT, OUT, 3, dtype=torch.uint8, device=dev)
    hdr = d(plan.hdr)
    if plan.geom is not None:
        hb, hk, vb, vk = ops.resample_taps(d(plan.geom), OUT, plan.ks_h, plan.ks_v, hdr)
    else:
        hb, hk, vb, vk = d(plan.hb), d(plan.hk), d(plan.vb), d(plan.vk)
    ops.resample_u8(src, hdr, hb, hk, vb, vk, OUT, OUT, tmp, x_orig[:Vr])
    if plan.x_pre is not None and V > 1:
        x_orig[1:].copy_(torch.from_numpy(plan.x_pre), non_blocking=True)
    if out is None:
        out = torch.empty(V, 3, OUT, OUT, dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (V, 3, OUT, OUT) or out.dtype != torch.float32 or not out.is_contiguous():
        raise RlcfError(f"run_plan: out must be a contiguous fp32 [{V},3,{OUT},{OUT}] tensor")
    ops.augmix_views(x_orig, d(plan.vflag), d(plan.wts), d(plan.omm), d(plan.n_ops), d(plan.ops), d(plan.mats), MEAN, STD,
                     out)
    return out


class AugMixAugmenter:
    """datautils.AugMixAugmenter (datautils.py:114-128) on the GPU.  `base_transform` / `preprocess` are accepted for
    signature compatibility; the fixed pipeline of tune_cls_rl.py:102-108 (Resize 224 bicubic + CenterCrop, ToTensor +
    CLIP normalisation) is what the kernels implement.  With hard_aug the BYOL pre-augmentation (colour jitter, grayscale,
    blur: PIL work of the data loader) runs on the host with the reference's torchvision calls; the rest is unchanged."""

    def __init__(self, base_transform=None, preprocess=None, n_views=2, augmix=False, severity=1, hard_aug=False,
                 device="cuda"):
        self.n_views, self.augmix, self.severity, self.device = n_views, bool(augmix), severity, device
        self.hard_aug = bool(hard_aug)

    def views(self, x) -> torch.Tensor:
        img = _to_u8_hwc(x)
        if self.hard_aug:
            plan = sample_plan_hard(x, self.n_views, self.augmix, self.severity, host_taps=False)
        else:
            plan = sample_plan(img.shape[1], img.shape[0], self.n_views, self.augmix, self.severity, host_taps=False)
        return run_plan(img, plan, self.device)

    def __call__(self, x):
        return list(self.views(x))
