"""Bring-up probe: watch one buffer (run.dres16) across every op of a 2-step retrieval run and name the op that
changes it unexpectedly."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from rlcf_b200 import ops, _lib
from test_oracle_retrieval import load_case, retrieval_setup
from test_retrieval_gpu import build, DEV

name = "ret_i2t_tiny_recipe"
z, cfg = load_case(name)
sd_p, sd_r, images, tokens, rcfg, gal_p, gal_r = retrieval_setup(cfg)
rcfg.tta_steps = 2
nq = cfg["n_query"]
eng = build(cfg, rcfg, sd_p, sd_r, gal_p, gal_r, nq)
q = images[:nq].to(DEV)
watch = {"dres16": eng.run.dres16, "g16": eng.run.g16, "dres": eng.run.dres}
snap = {k: v.clone() for k, v in watch.items()}
orig_call = _lib.call
n = [0]


def call(name, *args):
    orig_call(name, *args)
    torch.cuda.synchronize()
    n[0] += 1
    ptrs = {int(a) for a in args if isinstance(a, int) and a > (1 << 32)}
    for k, v in watch.items():
        changed = not torch.equal(v.view(torch.int16 if v.dtype == torch.float16 else torch.int32),
                                  snap[k].view(torch.int16 if v.dtype == torch.float16 else torch.int32))
        if changed:
            lo, hi = v.data_ptr(), v.data_ptr() + v.numel() * v.element_size()
            targeted = any(lo <= p < hi for p in ptrs)
            diff = (v.float() != snap[k].float()) | (torch.isnan(v.float()) != torch.isnan(snap[k].float()))
            rows = diff.any(dim=1).nonzero().flatten().tolist()
            print(f"op#{n[0]} {name}: {k} changed rows {rows[:6]}..{rows[-1] if rows else ''} ({int(diff.sum())} elems) "
                  f"{'(an argument points into it)' if targeted else '<-- NOT AN ARGUMENT'} nonfinite={int((~torch.isfinite(v.float())).sum())}")
            snap[k].copy_(v)
            if not targeted:
                print("   args:", [hex(a) if isinstance(a, int) and a > (1 << 32) else a for a in args])
                for nm, t in (("dres16", eng.run.dres16), ("dx_pre", eng.hook.dx_pre), ("partials", eng.partials),
                              ("x_pre", eng.store.x_pre), ("ln", eng.ln), ("dres", eng.run.dres), ("g16", eng.run.g16),
                              ("hook.y", eng.hook.y), ("hook.df", eng.hook.df), ("t_dy", eng.hook.t_dy)):
                    print(f"   {nm}: {hex(t.data_ptr())} .. {hex(t.data_ptr() + t.numel() * t.element_size())} shape {tuple(t.shape)}")
                first = diff.nonzero()[0].tolist()
                print("   first changed elem", first, "byte addr", hex(v.data_ptr() + (first[0] * v.shape[1] + first[1]) * v.element_size()))


_lib.call = call
ops.call = call
eng.tune(q)
print("done", n[0], "ops")
