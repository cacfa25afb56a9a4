"""Time and error against an fp32 torch reference for the attention kernel and the c_fc (QuickGELU epilogue) GEMM at the
hot-path shapes of a 32-image step.  Used for the round-1 A/B runs summarised in profiles/r1_attention_probes.txt."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
tag = f"attn_debug={os.environ.get('RLCF_ATTN_DEBUG', '0')}"


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


if "attn" in sys.argv[1:] or len(sys.argv) == 1:
    for (n_seq, L, heads) in [(2048, 197, 12), (192, 257, 16)]:
        d = heads * 64
        qkv = (torch.randn(n_seq * L, 3 * d, device=dev) * 1.5).half()
        out = torch.empty(n_seq * L, d, device=dev, dtype=torch.float16)
        us = timeit(lambda: ops.attention_fwd(qkv, n_seq, L, heads, out))
        ns = 8   # error on the first sequences
        q, k, v = (qkv[: ns * L].float().view(ns, L, 3, heads, 64).permute(2, 0, 3, 1, 4))
        ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v
        got = out[: ns * L].float().view(ns, L, heads, 64).permute(0, 2, 1, 3)
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        print(f"[{tag}] attention n_seq {n_seq} L {L}: {us:.1f} us  {4.0*L*L*64*heads*n_seq/us/1e6:.1f} TFLOP/s  max err/scale {err:.2e}", flush=True)

if "gelu" in sys.argv[1:] or len(sys.argv) == 1:
    for name, M, N, K in [("c_fc policy", 32 * 64 * 197, 3072, 768), ("c_fc reward", 32 * 6 * 257, 4096, 1024)]:
        a = torch.randn(M, K, device=dev).half()
        b = (torch.randn(N, K, device=dev) * 0.05).half()
        bias = torch.randn(N, device=dev)
        out = torch.zeros(M, N, device=dev, dtype=torch.float16)
        us = timeit(lambda: ops.gemm(a, b, out, epilogue=ops.EPI_GELU_F16, bias=bias))
        u = a[:4096].float() @ b.float().t() + bias
        ref = u * torch.sigmoid(1.702 * u)
        err = (out[:4096].float() - ref).abs().max().item() / ref.abs().max().item()
        print(f"[{tag}] {name} {M}x{N}x{K}: {us:.1f} us  {2.0*M*N*K/us/1e6:.1f} TFLOP/s  max err/scale {err:.2e}", flush=True)
