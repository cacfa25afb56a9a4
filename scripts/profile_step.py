"""One eager RLCF step (config 2, B images) between cudaProfilerStart/Stop -- the command ncu wraps:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python scripts/profile_step.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import engine as E, synthetic as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
sd_p, sd_r = S.make_state_dict("ViT-B/16", 0, dev), S.make_state_dict("ViT-L/14", 1, dev)
pol, rew = E.prepare_visual(sd_p, need_grad=True), E.prepare_visual(sd_r)
tok = S.make_tokens(200, 49408)
cf, rc = E.text_features(E.prepare_text(sd_p), tok), E.text_features(E.prepare_text(sd_r), tok)
cfg = E.RlcfConfig(n_views=64, selection_p=0.1, tta_steps=1, sample_k=3, lr=5e-3)
eng = E.RlcfEngine(pol, cf, float(sd_p["logit_scale"].exp()), cfg, B, reward=rew, reward_class_feat=rc)
views = S.make_views(B, 64, 224, 3, device=dev)
for _ in range(2):
    eng.adapt(views)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.adapt(views)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
