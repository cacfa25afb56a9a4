import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops, _lib
dev = torch.device("cuda:0")
n_seq, L, heads = 192, 197, 12
d = heads * 64
qkv = torch.randn(n_seq * L, 3 * d, device=dev).half()
dout = (torch.randn(n_seq * L, d, device=dev) * 0.1).half()
out = torch.empty(n_seq * L, d, device=dev, dtype=torch.float16)
lse = torch.empty(n_seq, heads, L, device=dev)
dqkv = torch.empty(n_seq * L, 3 * d, device=dev, dtype=torch.float16)
ops.attention_fwd(qkv, n_seq, L, heads, out, lse=lse)
ops.attention_bwd(qkv, out, dout, lse, n_seq, L, heads, dqkv)
torch.cuda.synchronize()
