"""CPU-side checks of the retrieval driver (rlcf_b200/clip_ret_policy.py): recall metrics against the values the
reference's RetrievalTask._report_metrics produced (committed fixtures), the query sharding, and a world_size-2 gloo
run of the score-row all-gather."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from test_oracle_retrieval import CASES, load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", CASES)
def test_report_metrics_match_reference(name):
    from rlcf_b200.clip_ret_policy import report_metrics
    z, _ = load_case(name)
    m = report_metrics(z["metrics_s_i2t"], z["metrics_s_t2i"], z["metrics_txt2img"].tolist(), z["metrics_img2txt"].tolist())
    assert sorted(m) == z["metrics_keys"].tolist()
    assert np.array_equal(np.array([m[k] for k in sorted(m)]), z["metrics_vals"])


def test_report_metrics_hand_case():
    from rlcf_b200.clip_ret_policy import report_metrics
    # 2 images x 4 captions; image 0 owns captions {0, 1}, image 1 owns {2, 3}
    s_i2t = np.array([[0.9, 0.1, 0.5, 0.2], [0.3, 0.8, 0.1, 0.7]])
    s_t2i = s_i2t.T.copy()
    m = report_metrics(s_i2t, s_t2i, [0, 0, 1, 1], [[0, 1], [2, 3]])
    assert m["txt_r1"] == 50.0 and m["txt_r5"] == 100.0          # image 1's best caption (3) is ranked second
    assert m["img_r1"] == 50.0                                   # captions 0 and 3 retrieve their image first
    assert m["agg_metrics"] == m["txt_r_mean"]


def test_query_sharding_covers_everything_once():
    from rlcf_b200.clip_ret_policy import shard_indices
    for n, w in ((10, 3), (5, 8), (25000, 8)):
        got = sorted(i for r in range(w) for i in shard_indices(n, r, w))
        assert got == list(range(n))


GLOO_WORKER = r"""
import sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from rlcf_b200.clip_ret_policy import gather_score_rows, shard_indices
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
n, C = 7, 5
full = torch.arange(n * C, dtype=torch.float32).view(n, C) * 0.5
mine = full[shard_indices(n, r, w)]
out = gather_score_rows(mine, n, r, w)
assert torch.equal(out, full), (r, out)
if r == 0:
    print("GATHER OK")
dist.destroy_process_group()
"""


def test_two_rank_gloo_score_row_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29547", str(script), ROOT],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER OK" in out.stdout
