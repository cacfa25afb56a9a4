"""Attention forward probes at the policy shape of a 32-image step (2048 sequences x 197 tokens x 12 heads):
RLCF_ATTN_DEBUG decomposition (1 skip max, 2 skip exp, 4 skip stores, 16 no exp token, 32 per-thread O stores) and, with
bit 8, a clock64 timeline of the first tiles of CTA 0 (both teams)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcf_b200 import ops
dev = torch.device("cuda:0")
dbg = int(os.environ.get("RLCF_ATTN_DEBUG", "0"))
n_seq, L, heads = 2048, 197, 12
d = heads * 64
qkv = (torch.randn(n_seq * L, 3 * d, device=dev) * 1.5).half()
out = torch.empty(n_seq * L, d, device=dev, dtype=torch.float16)
lse = torch.zeros(n_seq * heads * L, device=dev, dtype=torch.float32) if dbg & 8 else None
for _ in range(3):
    ops.attention_fwd(qkv, n_seq, L, heads, out, lse=lse)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention_fwd(qkv, n_seq, L, heads, out, lse=lse)
e1.record(); torch.cuda.synchronize()
print(f"[debug={dbg}] {e0.elapsed_time(e1)/10*1e3:.1f} us", flush=True)
if dbg & 8:
    st = lse.view(torch.int64)[: 2 * 64 * 8].view(2, 64, 8).cpu()
    t0 = int(st[:, 0, 0].min())
    names = ["start", "S ready", "max done", "P written", "O ready", "stored"]
    for team in range(2):
        print(f"team {team}: cycles since the first stamp; columns = {names}; then durations")
        for tc in range(12, 24):
            r = [int(x) - t0 for x in st[team, tc, :6]]
            dur = [r[i + 1] - r[i] for i in range(5)]
            print(f"  tile {tc:2d}: {r}  d={dur}  tile-to-tile {int(st[team, tc, 0]) - int(st[team, tc - 1, 0])}")
