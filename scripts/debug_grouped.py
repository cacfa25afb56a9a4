"""Bring-up probe: wraps ops.gemm_grouped and checks every grouped launch of a retrieval run against torch.matmul."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from rlcf_b200 import ops
from test_oracle_retrieval import load_case, retrieval_setup
from test_retrieval_gpu import build, DEV

orig = ops.gemm_grouped
count = [0]


def checked(a, b, out, epilogue=ops.EPI_F16, bias=None, resid=None, aux_in=None, aux_out=None, alpha=1.0):
    resid_c = None if resid is None else resid.clone()
    r = orig(a, b, out, epilogue=epilogue, bias=bias, resid=resid, aux_in=aux_in, aux_out=aux_out, alpha=alpha)
    torch.cuda.synchronize()
    ref = torch.matmul(a.float(), b.float().transpose(1, 2)) * alpha
    if bias is not None:
        ref = ref + bias[:, None, :]
    if epilogue == ops.EPI_RESID_F32:
        ref = ref + resid_c
    if epilogue == ops.EPI_GELU_F16:
        ref = ref * torch.sigmoid(1.702 * ref)
    if epilogue == ops.EPI_GELU_BWD_F16:
        u = aux_in.float(); s = torch.sigmoid(1.702 * u)
        ref = ref * (s * (1 + 1.702 * u * (1 - s)))
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    count[0] += 1
    flag = "" if err <= 2e-2 * max(scale, 1e-6) and torch.isfinite(out.float()).all() else "   <-- BAD"
    print(f"#{count[0]} G={a.shape[0]} m={a.shape[1]} n={b.shape[1]} k={a.shape[2]} epi={epilogue} "
          f"a.stride={a.stride()} b.stride={b.stride()} out.stride={out.stride()} err={err:.3e} scale={scale:.3e}{flag}")
    return r


ops.gemm_grouped = checked


def fin(t):
    return None if t is None else int((~torch.isfinite(t.float())).sum().item())


def wrap(name, outs):
    f = getattr(ops, name)

    def g(*a, **k):
        if name == "layernorm_bwd" and k.get("dx16") is not None:
            saved = (k["dx"].clone(), k["dx16"].clone(), a[6].clone())
        pre = {i: fin(t) for i, t in enumerate(a) if isinstance(t, torch.Tensor)}
        pre.update({kk: fin(t) for kk, t in k.items() if isinstance(t, torch.Tensor)})
        r = f(*a, **k)
        torch.cuda.synchronize()
        post = {i: fin(t) for i, t in enumerate(a) if isinstance(t, torch.Tensor)}
        post.update({kk: fin(t) for kk, t in k.items() if isinstance(t, torch.Tensor)})
        if any(v for v in post.values()) or any(v for v in pre.values()):
            print(f"   {name}: nonfinite before {pre} after {post}")
        if name == "layernorm_bwd" and k.get("dx16") is not None:
            dx, dx16 = k["dx"], k["dx16"]
            n = min(dx.shape[0], dx16.shape[0])
            bad = ~torch.isfinite(dx16[:n].float())
            if bad.any():
                for rep in range(3):
                    k["dx"].copy_(saved[0]); k["dx16"].copy_(saved[1]); a[6].copy_(saved[2])
                    torch.cuda.synchronize()
                    f(*a, **k)
                    torch.cuda.synchronize()
                    b2 = ~torch.isfinite(k["dx16"].float())
                    print("   rerun", rep, "bad", b2.nonzero()[:6].tolist(), "dx finite", bool(torch.isfinite(k["dx"]).all()),
                          "dy finite", bool(torch.isfinite(a[0].float()).all()), "ptrs", hex(k["dx"].data_ptr()), hex(k["dx16"].data_ptr()),
                          hex(a[0].data_ptr()), hex(a[1].data_ptr()), hex(a[6].data_ptr()), a[6].shape)
                idx = bad.nonzero()
                print("   dx16 bad at", idx[:8].tolist(), "dx there", dx[:n][bad][:8].tolist(), "dx16", dx16[:n][bad][:8].tolist(),
                      "shapes", tuple(dx.shape), tuple(dx16.shape), "rows arg", a[3], a[4])
        if name == "layernorm_bwd":
            am = lambda t: float(t.float().abs().nan_to_num(0, 0, 0).max())
            x = a[1]
            print(f"   ln_bwd: |dy|={am(a[0]):.3e} |x|={am(x):.3e} minstd={float(x.std(dim=-1).min()):.3e} "
                  f"|gamma|={am(a[2][:x.shape[1]]):.3e} |dx|={am(k['dx']) if k.get('dx') is not None else -1:.3e} p_off={a[9]}")
        return r
    setattr(ops, name, g)


for nm in ("layernorm_bwd", "layernorm_fwd", "attention_bwd", "attention_fwd", "head_bwd_ex", "adamw_step",
           "adamw_step_from", "transpose_blocks", "colsum_f16", "seq_sum", "outer_sum", "embed_lnpre", "head_fwd"):
    wrap(nm, None)
name = sys.argv[1] if len(sys.argv) > 1 else "ret_i2t_tiny_recipe"
z, cfg = load_case(name)
sd_p, sd_r, images, tokens, rcfg, gal_p, gal_r = retrieval_setup(cfg)
rcfg.tta_steps = 2
nq = cfg["n_query"]
eng = build(cfg, rcfg, sd_p, sd_r, gal_p, gal_r, nq)
q = (images if cfg["task"] == "image2text" else tokens)[:nq].to(DEV)
eng.tune(q)
torch.cuda.synchronize()
print("done")
