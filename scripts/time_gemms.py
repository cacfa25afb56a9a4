"""Times the tcgen05 GEMM at the hot-path shapes of one B-image step (policy B/16 and reward L/14)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
Mp, Mr = B * 64 * 197, B * 6 * 257
shapes = [("qkv   policy", Mp, 2304, 768, ops.EPI_F16), ("oproj policy", Mp, 768, 768, ops.EPI_RESID_F32),
          ("c_fc  policy", Mp, 3072, 768, ops.EPI_GELU_F16), ("cproj policy", Mp, 768, 3072, ops.EPI_RESID_F32),
          ("qkv   reward", Mr, 3072, 1024, ops.EPI_F16), ("oproj reward", Mr, 1024, 1024, ops.EPI_RESID_F32),
          ("c_fc  reward", Mr, 4096, 1024, ops.EPI_GELU_F16), ("cproj reward", Mr, 1024, 4096, ops.EPI_RESID_F32)]
tot = 0.0
for name, M, N, K, epi in shapes:
    a = torch.randn(M, K, device=dev).half()
    b = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    f32 = epi in (ops.EPI_RESID_F32, ops.EPI_F32)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.float16)
    kw = dict(epilogue=epi, bias=bias)
    if epi == ops.EPI_RESID_F32:
        kw["resid"] = out
    for _ in range(2):
        ops.gemm(a, b, out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        ops.gemm(a, b, out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    layers = 12 if "policy" in name else 24
    tot += ms * layers
    print(f"{name} {M}x{N}x{K} epi{epi}: {ms*1e3:7.1f} us {2*M*N*K/ms/1e9:7.1f} TFLOP/s   x{layers} layers = {ms*layers:.2f} ms", flush=True)
print(f"sum over layers {tot:.2f} ms")
