"""Diagnostic: where does the logit error of the CUDA path come from? (image tower vs text tower)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import rlcf_oracle as O
from rlcf_b200 import engine as E

torch.set_num_threads(os.cpu_count())
arch = sys.argv[1] if len(sys.argv) > 1 else "ViT-B/16"
sd = O.make_clip_state_dict(arch, 0)
tok = O.make_tokens(200, O.ARCHS[arch][6], seed=7)
cf = O.class_features(sd, tok)
views = O.make_views(1, 64, O.ARCHS[arch][1], 11)[:16]
with torch.no_grad():
    f = O.encode_image(sd, views)
    f = f / f.norm(dim=-1, keepdim=True)
    ref = 100 * f @ cf.t()
sdd = {k: v.cuda() for k, v in sd.items()}
gi = E.image_features(E.prepare_visual(sdd), views.cuda()).cpu()
gt = E.text_features(E.prepare_text(sdd), tok).cpu()
print("image feat rel err (mean over views):", ((gi - f).norm(dim=-1) / f.norm(dim=-1)).mean().item())
print("text  feat rel err (mean over classes):", ((gt - cf).norm(dim=-1) / cf.norm(dim=-1)).mean().item())
sc = ref.abs().max()
for name, lg in (("cuda img x oracle txt", 100 * gi @ cf.t()), ("oracle img x cuda txt", 100 * f @ gt.t()),
                 ("cuda img x cuda txt", 100 * gi @ gt.t())):
    print(f"{name}: max rel {((lg - ref).abs().max() / sc).item():.3e} rms rel {((lg - ref).pow(2).mean().sqrt() / sc).item():.3e}")
