"""GPU bring-up: correctness + timing of the tcgen05 GEMM for one CTA-group mode (run under `timeout`)."""
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlcf_b200 import _lib, ops

cg = int(sys.argv[1])
print("cta_group", _lib.set_gemm_cta_group(cg), flush=True)
dev = torch.device("cuda:0")
torch.manual_seed(0)
ok = True
for (M, N, K) in [(128, 256, 64), (256, 256, 128), (300, 768, 768), (1000, 32, 64), (12608, 2304, 768)]:
    a = torch.randn(M, K, device=dev).half()
    b = torch.randn(N, K, device=dev).half()
    out = torch.zeros(M, N, device=dev, dtype=torch.float32)
    ops.gemm(a, b, out, epilogue=ops.EPI_F32)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    print(f"shape {M}x{N}x{K}: rel err {err:.3e}", flush=True)
    if not err < 2e-5:
        ok = False
        bad = (out - ref).abs() > 1e-2 * ref.abs().max()
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("  bad rows", rows[:8].tolist(), "...", rows[-4:].tolist(), "n", rows.numel(),
              "bad cols", cols[:8].tolist(), "...", cols[-4:].tolist(), "n", cols.numel(), flush=True)
        print("  out[0,:8]", out[0, :8].tolist(), "ref[0,:8]", ref[0, :8].tolist(), flush=True)
print("CORRECT" if ok else "WRONG", flush=True)

# timing on the hot-path shapes (inputs larger than nothing special: L2-resident, this is a kernel-speed probe)
for (M, N, K, epi) in [(12608, 2304, 768, ops.EPI_F16), (12608, 768, 768, ops.EPI_RESID_F32),
                       (12608, 3072, 768, ops.EPI_GELU_F16), (12608, 768, 3072, ops.EPI_RESID_F32),
                       (50432, 2304, 768, ops.EPI_F16), (50432, 768, 3072, ops.EPI_RESID_F32),
                       (8192, 8192, 8192, ops.EPI_F16)]:
    a = torch.randn(M, K, device=dev).half()
    b = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    f32 = epi in (ops.EPI_RESID_F32, ops.EPI_F32)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.float16)
    kw = dict(epilogue=epi, bias=bias)
    if epi == ops.EPI_RESID_F32:
        kw["resid"] = out
    for _ in range(3):
        ops.gemm(a, b, out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ops.gemm(a, b, out, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"time {M}x{N}x{K} epi{epi}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    # cuBLAS reference speed for the same shape
    c = torch.empty(M, N, device=dev, dtype=torch.float16)
    for _ in range(3):
        torch.matmul(a, b.t(), out=c)
    e0.record()
    for _ in range(n):
        torch.matmul(a, b.t(), out=c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"   cublas fp16 same shape: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
