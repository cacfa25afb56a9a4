"""Launches the long-sequence / row attention kernels once each at their hot shapes (for `ncu --set full -k regex:...`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops
dev = torch.device("cuda:0")
for (n_seq, L, heads) in [(384, 577, 16), (384, 257, 16)]:
    d = heads * 64
    qkv = torch.randn(n_seq * L, 3 * d, device=dev).half()
    out = torch.empty(n_seq * L, d, device=dev, dtype=torch.float16)
    for _ in range(3):
        ops.attention_fwd(qkv, n_seq, L, heads, out)
n_seq, L, heads = 4096, 197, 12
d = heads * 64
qkv = torch.randn(n_seq * L, 3 * d, device=dev).half()
x = torch.randn(n_seq * L, d, device=dev)
out = torch.empty(n_seq, d, device=dev, dtype=torch.float16)
xr = torch.empty(n_seq, d, device=dev)
for _ in range(3):
    ops.attention_row_fwd(qkv, n_seq, L, heads, out, x=x, x_row=xr)
torch.cuda.synchronize()
