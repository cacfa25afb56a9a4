"""Throughput of the on-device view generation (64 views of one 500x375 image) next to the reference's PIL pipeline on
one host thread.  Reported per image: host plan (random decisions + Pillow tap tables), upload + kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import augmix_oracle as A          # the PIL pipeline, as the timed CPU baseline only
from rlcf_b200 import datautils as D

dev = torch.device("cuda:0")
img = A.synthetic_image(375, 500, 3)
u8 = D._to_u8_hwc(img)
for augmix in (False, True):
    torch.manual_seed(0); np.random.seed(0)
    aug = D.AugMixAugmenter(n_views=63, augmix=augmix, device=dev)
    for _ in range(3):
        aug.views(img)
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    plans = [D.sample_plan(500, 375, 63, augmix, host_taps=False) for _ in range(n)]
    t_plan = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for p in plans:
        D.run_plan(u8, p, dev)
    e1.record()
    torch.cuda.synchronize()
    t_run = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(3):
        A.augmix_views(img, 63, augmix)
    t_pil = (time.perf_counter() - t0) / 3
    print(f"augmix={augmix}: host plan {t_plan*1e3:.2f} ms/img, upload+kernels {t_run*1e3:.2f} ms/img (wall) "
          f"{e0.elapsed_time(e1)/n:.2f} ms/img (device) -> {1/(t_plan+t_run):.0f} img/s single-threaded; "
          f"PIL pipeline (1 thread) {t_pil*1e3:.1f} ms/img = {1/t_pil:.1f} img/s; "
          f"bytes uploaded per image {u8.numel() + sum(a.nbytes for a in (p.hdr,p.geom,p.vflag,p.wts,p.omm,p.n_ops,p.ops,p.mats))} "
          f"vs {64*3*224*224*4} of fp32 views", flush=True)


# ---- the adaptation fed from uint8 images: views generated on the device straight into the engine's input batch
from rlcf_b200 import engine as E, synthetic as S
B, V = 32, 64
sd_p, sd_r = S.make_state_dict("ViT-B/16", 0, dev), S.make_state_dict("ViT-L/14", 1, dev)
tok = S.make_tokens(200, 49408)
cf, rcf = E.text_features(E.prepare_text(sd_p), tok), E.text_features(E.prepare_text(sd_r), tok)
eng = E.RlcfEngine(E.prepare_visual(sd_p, need_grad=True), cf, float(sd_p["logit_scale"].exp()),
                   E.RlcfConfig(n_views=V, selection_p=0.1, tta_steps=1, sample_k=3, lr=5e-3), B,
                   reward=E.prepare_visual(sd_r), reward_class_feat=rcf)
del sd_p, sd_r
batch = torch.empty(B * V, 3, 224, 224, device=dev)
eng.capture(batch)
imgs = [D._to_u8_hwc(A.synthetic_image(375, 500, 100 + i)).pin_memory() for i in range(B)]
torch.manual_seed(0); np.random.seed(0)
plans = [D.sample_plan(500, 375, V - 1, False, host_taps=False) for _ in range(B)]   # host work of the loader workers
out_host = torch.empty(B, 200).pin_memory()


def step():
    for i in range(B):
        D.run_plan(imgs[i], plans[i], dev, out=eng._static_images[i * V:(i + 1) * V])
    eng._graph.replay()
    out_host.copy_(eng.logits_final, non_blocking=True)


t0 = time.perf_counter()
while time.perf_counter() - t0 < 3.0:
    step()
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 8
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
h2d = sum(i.numel() for i in imgs) + sum(sum(a.nbytes for a in (p.hdr, p.geom, p.vflag, p.wts, p.omm, p.n_ops, p.ops, p.mats)) for p in plans)
print(f"adaptation fed from uint8 images (views generated on the device, plans precomputed): {B / ms * 1e3:.1f} images/s, "
      f"{ms:.1f} ms per {B}-image step, H2D {h2d / 1e6:.1f} MB per step vs {B * V * 3 * 224 * 224 * 4 / 1e6:.0f} MB of fp32 views", flush=True)
