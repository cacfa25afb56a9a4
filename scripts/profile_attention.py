"""A few launches of the tcgen05 attention forward at the policy shape (the command `ncu --set full` wraps):
   ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 3 -c 1 -o X python scripts/profile_attention.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops
n_seq, L, heads = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (512, 197, 12)))
dev = torch.device("cuda:0")
d = heads * 64
qkv = (torch.randn(n_seq * L, 3 * d, device=dev) * 1.5).half()
out = torch.empty(n_seq * L, d, device=dev, dtype=torch.float16)
for _ in range(5):
    ops.attention_fwd(qkv, n_seq, L, heads, out)
torch.cuda.synchronize()
