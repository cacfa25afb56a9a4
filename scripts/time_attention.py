"""Times the forward attention kernel alone at the two hot-path shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops, _lib
dev = torch.device("cuda:0")
for impl in ([0, 1] if len(sys.argv) < 2 else [int(sys.argv[1])]):
    _lib.set_attention_impl(impl)
    for (n_seq, L, heads) in [(512, 197, 12), (48, 257, 16), (384, 257, 16), (48, 197, 12), (48, 577, 16), (384, 577, 16)]:
        d = heads * 64
        qkv = torch.randn(n_seq * L, 3 * d, device=dev).half()
        out = torch.empty(n_seq * L, d, device=dev, dtype=torch.float16)
        for _ in range(3):
            ops.attention_fwd(qkv, n_seq, L, heads, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.attention_fwd(qkv, n_seq, L, heads, out)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        fl = 4.0 * L * L * 64 * heads * n_seq
        tiles = ((L + 127) // 128) * heads * n_seq
        print(f"impl {impl} n_seq {n_seq} L {L} heads {heads}: {us:.1f} us  {fl/us/1e6:.1f} TFLOP/s  {us*1e3/ (tiles/148):.0f} ns per tile-slot", flush=True)
