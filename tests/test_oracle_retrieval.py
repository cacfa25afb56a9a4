"""Pins the oracle's retrieval restatement (oracle/rlcf_oracle.py: retrieval_*) to outputs of the reference's own
retrieval TTA (tests/golden/ret_*.npz, made by oracle/make_golden_retrieval.py from /root/reference/retrieval).
Runs without a GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import rlcf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POLICY_SEED, REWARD_SEED, IMAGE_SEED, TOKEN_SEED = 0, 1, 21, 9
CASES = ["ret_i2t_tiny_3step", "ret_i2t_tiny_recipe", "ret_t2i_tiny_3step", "ret_t2i_tiny_recipe"]


def load_case(name):
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    z = np.load(path)
    cfg = dict(zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist()))
    for k in ("n_query", "n_gallery", "K", "steps", "momentum_update", "update_freq"):
        if k in cfg:
            cfg[k] = int(cfg[k])
    for k in ("lr", "update_w", "momentum"):
        if k in cfg:
            cfg[k] = float(cfg[k])
    return z, cfg


def retrieval_setup(cfg):
    """Weights, inputs and gallery features of a golden case, the way clip_ret_policy.test_time_tune prepares them."""
    i2t = cfg["task"] == "image2text"
    sd_p = O.openai_load_rounding(O.make_clip_state_dict(cfg["policy"], POLICY_SEED))
    sd_r = O.openai_load_rounding(O.make_clip_state_dict(cfg["reward"], REWARD_SEED))
    nq, ng = cfg["n_query"], cfg["n_gallery"]
    n_img, n_txt = (nq, ng) if i2t else (ng, nq)
    images = O.make_views(n_img, 1, O.ARCHS[cfg["policy"]][1], IMAGE_SEED)
    tokens = O.make_tokens(n_txt, O.ARCHS[cfg["policy"]][6], seed=TOKEN_SEED)
    rcfg = O.RetrievalConfig(tta_steps=cfg["steps"], sample_k=cfg["K"], lr=cfg["lr"],
                             momentum_update=bool(cfg.get("momentum_update", 0)),
                             update_freq=cfg.get("update_freq", 256), update_w=cfg.get("update_w", 1.0),
                             momentum=cfg.get("momentum", 0.9999))
    with torch.no_grad():
        if i2t:
            gal_p, gal_r = O.retrieval_features(sd_p, tokens=tokens), O.retrieval_features(sd_r, tokens=tokens)
        else:
            gal_p, gal_r = O.retrieval_features(sd_p, images=images), O.retrieval_features(sd_r, images=images)
    return sd_p, sd_r, images, tokens, rcfg, gal_p, gal_r


@pytest.mark.parametrize("name", CASES)
def test_retrieval_oracle_matches_reference(name):
    z, cfg = load_case(name)
    torch.set_num_threads(os.cpu_count() or 1)
    i2t = cfg["task"] == "image2text"
    sd_p, sd_r, images, tokens, rcfg, gal_p, gal_r = retrieval_setup(cfg)
    assert np.abs(gal_p.numpy() - z["gallery_policy"]).max() < 1e-5
    assert np.abs(gal_r.numpy() - z["gallery_reward"]).max() < 1e-5
    state = O.RetrievalMomentum(sd_p, rcfg)
    stride = int(z["param_stride"])
    rows = []
    for q in range(cfg["n_query"]):
        query = images[q:q + 1] if i2t else tokens[q:q + 1]
        with torch.no_grad():
            rq = O.retrieval_features(sd_r, images=query) if i2t else O.retrieval_features(sd_r, tokens=query)
        assert np.abs(rq.numpy() - z[f"q{q}.reward_query"]).max() < 1e-5
        out = O.retrieval_tune_query(state.initial, rcfg, cfg["task"], query, gal_p, rq, gal_r)
        assert np.array_equal(torch.stack(out["topk_idx"]).numpy(), z[f"q{q}.topk_idx"])
        assert np.abs(torch.stack(out["scores"]).numpy() - z[f"q{q}.scores"]).max() < 1e-5
        assert np.abs(torch.stack(out["rewards"]).numpy() - z[f"q{q}.rewards"]).max() < 1e-5
        row = z[f"q{q}.score_row"]
        assert np.abs(out["score_row"].numpy() - row).max() < 1e-4 * np.abs(row).max()
        # adapted parameters: AdamW moves a parameter by ~lr per step whatever the gradient's size, so two fp32 runs
        # may differ by a full step where the gradient is at rounding level; everywhere else they agree closely
        got = torch.cat([out["state"][n].flatten() for n in out["param_names"]])[::stride].numpy()
        want = z[f"q{q}.params"]
        err = np.abs(got - want)
        assert err.max() <= 2.05 * cfg["lr"] * cfg["steps"]
        assert np.mean(err > 0.05 * cfg["lr"]) < 0.02, np.mean(err > 0.05 * cfg["lr"])
        state.update(out["state"])
        init = torch.cat([state.initial[n].flatten() for n in out["param_names"]])[::stride].numpy()
        assert np.abs(init - z[f"q{q}.initial_after"]).max() <= 2.05 * cfg["lr"] * cfg["steps"]
        rows.append(out["score_row"].numpy())
    assert np.abs(np.stack(rows) - z["score_matrix"]).max() < 1e-4 * np.abs(z["score_matrix"]).max()


@pytest.mark.parametrize("name", CASES)
def test_recall_metrics_match_reference(name):
    z, _ = load_case(name)
    m = O.retrieval_report_metrics(z["metrics_s_i2t"], z["metrics_s_t2i"], z["metrics_txt2img"].tolist(),
                                   z["metrics_img2txt"].tolist())
    assert sorted(m) == z["metrics_keys"].tolist()
    assert np.array_equal(np.array([m[k] for k in sorted(m)]), z["metrics_vals"])
