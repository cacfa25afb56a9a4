"""A few launches of the tcgen05 GEMM at one hot-path shape (the command `ncu --set full` wraps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops

M, N, K, epi = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (12608, 2304, 768, 0))]
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev).half()
b = (torch.randn(N, K, device=dev) * 0.05).half()
bias = torch.randn(N, device=dev)
f32 = epi in (ops.EPI_RESID_F32, ops.EPI_F32)
out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.float16)
kw = dict(epilogue=epi, bias=bias)
if epi == ops.EPI_RESID_F32:
    kw["resid"] = out
for _ in range(4):
    ops.gemm(a, b, out, **kw)
torch.cuda.synchronize()
