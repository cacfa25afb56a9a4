"""Runs rlcf_layernorm_bwd on seeded inputs (hot-path geometry: 2 sets x 1182 rows, width 768 / 1024, 32 slots; fp16
and fp32 dy) and saves every output to argv[1]; with --time also prints the duration at the 32-image size.  The
library reads RLCF_LN_BWD_SMEM once per process, so the experimental variant is compared across two processes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops
dev = torch.device("cuda:0")
res = {}
for d in (768, 1024):
    for f32 in (False, True):
        torch.manual_seed(d + f32)
        n_sets, rows, n_slots = 2, 1182, 32
        M = n_sets * rows
        x = torch.randn(M, d, device=dev) * 2 + 0.5
        params = torch.randn(n_sets, 2 * d, device=dev)
        dy = torch.randn(M, d, device=dev)
        dy = dy if f32 else dy.half()
        partials = torch.zeros(n_sets, n_slots, 2 * d, device=dev)
        dres = torch.randn(M, d, device=dev)
        d16 = torch.empty(M, d, device=dev, dtype=torch.float16)
        ops.layernorm_bwd(dy, x, params, rows, n_sets, d, partials, n_slots, 2 * d, 0, dx=dres, accumulate=True,
                          param_stride=2 * d, dx16=d16)
        res[f"{d}_{int(f32)}"] = (partials.cpu(), dres.cpu(), d16.cpu())
torch.save(res, sys.argv[1])
if "--time" in sys.argv:
    d, n_sets, rows, n_slots = 768, 32, 1182, 32
    M = n_sets * rows
    x = torch.randn(M, d, device=dev); params = torch.randn(n_sets, 2 * d, device=dev)
    dy = torch.randn(M, d, device=dev).half(); partials = torch.zeros(n_sets, n_slots, 2 * d, device=dev)
    dres = torch.randn(M, d, device=dev); d16 = torch.empty(M, d, device=dev, dtype=torch.float16)
    f = lambda: ops.layernorm_bwd(dy, x, params, rows, n_sets, d, partials, n_slots, 2 * d, 0, dx=dres, accumulate=True,
                                  param_stride=2 * d, dx16=d16)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record(); torch.cuda.synchronize()
    print(f"ln_bwd smem={os.environ.get('RLCF_LN_BWD_SMEM', '0')}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
