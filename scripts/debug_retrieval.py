"""Bring-up probe for the retrieval engines: runs a golden case step by step and reports the first non-finite buffer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_oracle_retrieval import load_case, retrieval_setup
from test_retrieval_gpu import build, DEV

name = sys.argv[1] if len(sys.argv) > 1 else "ret_i2t_tiny_recipe"
z, cfg = load_case(name)
sd_p, sd_r, images, tokens, rcfg, gal_p, gal_r = retrieval_setup(cfg)
nq = cfg["n_query"]
eng = build(cfg, rcfg, sd_p, sd_r, gal_p, gal_r, nq)
i2t = cfg["task"] == "image2text"
q = (images if i2t else tokens)[:nq].to(DEV)


def report(tag):
    torch.cuda.synchronize()
    for nm in ("reward_feat", "feat", "logits", "dlogits", "df_partial", "grads", "rest", "ln", "w16", "score_rows"):
        t = getattr(eng, nm, None)
        if t is None:
            continue
        bad = (~torch.isfinite(t.float())).sum().item()
        print(f"{tag} {nm}: nonfinite={bad} absmax={t.float().abs().nan_to_num(0, 0, 0).max().item():.3e}")
    print(tag, "topk", eng.topk_idx[0].tolist())


if i2t:
    from rlcf_b200 import ops, full_tune as FT
    xr = eng.rrun.forward(nq, eng.reward.ln_flat, images=q)
    eng.rrun.head(xr, nq, eng.reward.ln_flat, feat=eng.reward_feat)
    ops.reset_params(eng.init_ln, eng.ln, eng.ln_m, eng.ln_v, nq, eng.base.P)
    for step in range(1, rcfg.tta_steps + 1):
        try:
            eng._step(step, q, eng.w0 if step == 1 else eng.gw)
            report(f"step{step}")
        except Exception as e:
            print("step", step, "failed:", e)
            report(f"step{step}-fail")
            break
else:
    eng.tune(q)
    report("tune")
