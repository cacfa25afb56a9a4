"""Where the c_fc GEMM loses against the qkv GEMM (same K = 768, same 256 x 256 tiles): epilogue variants at both shapes.
RLCF_GEMM_DEBUG_NOSTORE=1 drops everything after the TMEM load."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops
dev = torch.device("cuda:0")
M = 32 * 64 * 197
tag = "nostore" if os.environ.get("RLCF_GEMM_DEBUG_NOSTORE") else "full"
for name, N, K, epi in [("qkv/f16", 2304, 768, ops.EPI_F16), ("c_fc/f16", 3072, 768, ops.EPI_F16),
                        ("c_fc/gelu", 3072, 768, ops.EPI_GELU_F16), ("qkv/gelu", 2304, 768, ops.EPI_GELU_F16)]:
    a = torch.randn(M, K, device=dev).half()
    b = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.float16)
    for _ in range(3):
        ops.gemm(a, b, out, epilogue=epi, bias=bias)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(a, b, out, epilogue=epi, bias=bias)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    tiles = ((M + 255) // 256) * (N // 256)
    print(f"[{tag}] {name} N={N}: {us:.1f} us  {2.0*M*N*K/us/1e6:.1f} TFLOP/s  {us*1e3/tiles*74:.0f} ns per tile per CTA pair", flush=True)
