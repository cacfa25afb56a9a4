"""TEST INFRASTRUCTURE: a numpy executor of rlcf_b200.datautils.ViewPlan, statement for statement what
csrc/augment_kernels.cu does on the device.  Lets the CPU suite check the host-side plan (random decisions, Pillow taps,
affine coefficients) against the PIL oracle without a GPU."""
import numpy as np

from rlcf_b200 import datautils as D

PB = D.PRECISION_BITS


def clip8(v):
    return np.clip(v >> PB, 0, 255).astype(np.uint8)


def resample(src, plan):
    V = plan.hdr.shape[0]
    out = np.zeros((V, 224, 224, 3), dtype=np.uint8)
    for v in range(V):
        x0, y0, rf, nr, flip = (int(t) for t in plan.hdr[v, :5])
        tmp = np.zeros((nr, 224, 3), dtype=np.uint8)
        rows = src[y0 + rf:y0 + rf + nr].astype(np.int64)
        for x in range(224):
            xmin, cnt = int(plan.hb[v, x, 0]), int(plan.hb[v, x, 1])
            k = plan.hk[v, x, :cnt].astype(np.int64)
            s = (1 << (PB - 1)) + (rows[:, x0 + xmin:x0 + xmin + cnt, :] * k[None, :, None]).sum(axis=1)
            tmp[:, x, :] = clip8(s)
        t64 = tmp.astype(np.int64)
        for y in range(224):
            ymin, cnt = int(plan.vb[v, y, 0]), int(plan.vb[v, y, 1])
            k = plan.vk[v, y, :cnt].astype(np.int64)
            s = (1 << (PB - 1)) + (t64[ymin:ymin + cnt] * k[:, None, None]).sum(axis=0)
            out[v, y] = clip8(s)[::-1] if flip else clip8(s)
    return out


def affine(plane, a):
    ys, xs = np.mgrid[0:224, 0:224]
    xc, yc = xs + 0.5, ys + 0.5
    xin = a[0] * xc + a[1] * yc + a[2]
    yin = a[3] * xc + a[4] * yc + a[5]
    inside = ~((xin < 0.0) | (xin >= 224) | (yin < 0.0) | (yin >= 224))
    xin, yin = xin - 0.5, yin - 0.5
    xi, yi = np.floor(xin).astype(np.int64), np.floor(yin).astype(np.int64)
    dx, dy = xin - xi, yin - yi
    cl = lambda c: np.clip(c, 0, 223)   # noqa: E731
    p = plane.astype(np.int64)
    x0, x1 = cl(xi), cl(xi + 1)
    r0 = cl(yi)
    v1 = p[r0, x0] + (p[r0, x1] - p[r0, x0]) * dx
    has2 = (yi + 1 >= 0) & (yi + 1 < 224)
    r1 = cl(yi + 1)
    v2 = np.where(has2, p[r1, x0] + (p[r1, x1] - p[r1, x0]) * dx, v1)
    v = v1 + (v2 - v1) * dy
    return np.where(inside, v.astype(np.int64), 0).astype(np.uint8)


def lut_autocontrast(plane):
    h = np.bincount(plane.ravel(), minlength=256)
    nz = np.nonzero(h)[0]
    lo, hi = int(nz[0]), int(nz[-1])
    if hi <= lo:
        return plane
    scale = 255.0 / (hi - lo)
    offset = -lo * scale
    lut = np.array([min(255, max(0, int(i * scale + offset))) for i in range(256)], dtype=np.uint8)
    return lut[plane]


def lut_equalize(plane):
    h = np.bincount(plane.ravel(), minlength=256)
    histo = [int(c) for c in h if c]
    if len(histo) <= 1:
        return plane
    step = (sum(histo) - histo[-1]) // 255
    if not step:
        return plane
    n, lut = step // 2, []
    for i in range(256):
        lut.append(min(255, n // step))      # Image.point clips list entries to uint8 (CLIP8 in _imaging.c getlist)
        n += int(h[i])
    return np.array(lut, dtype=np.uint8)[plane]


def pre(plane, c):
    x = plane.astype(np.float32) / np.float32(255)
    return (x - np.float32(D.MEAN[c])) / np.float32(D.STD[c])


def augmix(x_orig, plan):
    V = x_orig.shape[0]
    out = np.zeros((V, 3, 224, 224), dtype=np.float32)
    for v in range(V):
        for c in range(3):
            src = x_orig[v, :, :, c]
            xp = pre(src, c)
            if plan.vflag[v] == 0:
                out[v, c] = xp
                continue
            mix = np.zeros((224, 224), dtype=np.float32)
            for ch in range(3):
                cur = src.copy()
                for k in range(int(plan.n_ops[v, ch])):
                    t, p = int(plan.ops[v, ch, k, 0]), int(plan.ops[v, ch, k, 1])
                    if t == D.OP_AFFINE:
                        cur = affine(cur, plan.mats[v, ch, k])
                    elif t == D.OP_AUTOCONTRAST:
                        cur = lut_autocontrast(cur)
                    elif t == D.OP_EQUALIZE:
                        cur = lut_equalize(cur)
                    elif t == D.OP_POSTERIZE:
                        cur = cur & np.uint8(~((1 << (8 - p)) - 1) & 0xFF)
                    elif t == D.OP_SOLARIZE:
                        ci = cur.astype(np.int64)
                        cur = np.where(ci < p, ci, 255 - ci).astype(np.uint8)
                mix = mix + plan.wts[v, ch] * pre(cur, c)
            out[v, c] = plan.wts[v, 3] * xp + plan.omm[v] * mix
    return out


def execute(image_u8, plan):
    x_orig = resample(image_u8, plan)
    if getattr(plan, "x_pre", None) is not None:        # hard_aug: views 1.. were pre-augmented on the host
        x_orig = np.concatenate([x_orig[:1], plan.x_pre], axis=0)
    return augmix(x_orig, plan)
