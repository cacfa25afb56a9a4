"""One eager retrieval adaptation (config 4 shapes, Q queries) between cudaProfilerStart/Stop -- the command ncu wraps:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X \
       python scripts/profile_retrieval.py i2t 8
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import engine as E, retrieval as R, synthetic as S

task = sys.argv[1] if len(sys.argv) > 1 else "i2t"
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
sd_p, sd_r = S.make_state_dict("ViT-B/16", 0, dev), S.make_state_dict("ViT-L/14", 1, dev)
n_gal = 25000 if task == "i2t" else 5000
g = torch.Generator(device=dev).manual_seed(5)
gal_p = torch.nn.functional.normalize(torch.randn(n_gal, 512, generator=g, device=dev), dim=-1)
gal_r = torch.nn.functional.normalize(torch.randn(n_gal, 768, generator=g, device=dev), dim=-1)
cfg = R.RetrievalConfig(tta_steps=steps, sample_k=20 if task == "i2t" else 12, lr=1e-6)
if task == "i2t":
    eng = R.ImageQueryEngine(sd_p, gal_p, float(sd_p["logit_scale"].exp()), cfg, Q, E.prepare_visual(sd_r), gal_r)
    q = S.make_views(Q, 1, 224, 3, device=dev)
else:
    eng = R.TextQueryEngine(sd_p, gal_p, cfg, Q, E.prepare_text(sd_r), gal_r)
    q = S.make_tokens(Q, 49408, seed=31).to(dev)
eng.adapt(q)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.adapt(q)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
