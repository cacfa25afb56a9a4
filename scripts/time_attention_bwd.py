"""Times the attention backward kernels alone at the hot-path shape (selected views of a 32-image step: 192 x 197 x 12)
and the text-tower shape of prompt tuning (1600 x 77 x 8, causal).  impl 0 = tcgen05, 1 = warp-MMA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops, _lib
dev = torch.device("cuda:0")
for (n_seq, L, heads, causal) in [(192, 197, 12, False), (1600, 77, 8, True), (48, 197, 12, False), (48, 257, 16, False)]:
    d = heads * 64
    qkv = torch.randn(n_seq * L, 3 * d, device=dev).half()
    dout = (torch.randn(n_seq * L, d, device=dev) * 0.1).half()
    out = torch.empty(n_seq * L, d, device=dev, dtype=torch.float16)
    lse = torch.empty(n_seq, heads, L, device=dev)
    dqkv = torch.empty(n_seq * L, 3 * d, device=dev, dtype=torch.float16)
    ops.attention_fwd(qkv, n_seq, L, heads, out, causal=causal, lse=lse)
    for impl in (0, 1):
        _lib.set_attention_impl(impl)
        for _ in range(3):
            ops.attention_bwd(qkv, out, dout, lse, n_seq, L, heads, dqkv, causal=causal)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.attention_bwd(qkv, out, dout, lse, n_seq, L, heads, dqkv, causal=causal)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        fl = 10.0 * L * L * 64 * heads * n_seq          # five L x L x 64 contractions (2 FLOP per MAC)
        print(f"impl {impl} n_seq {n_seq} L {L} heads {heads} causal {causal}: {us:.1f} us  {fl/us/1e6:.1f} TFLOP/s", flush=True)
    _lib.set_attention_impl(0)
