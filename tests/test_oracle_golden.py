"""Pins the CPU oracle (oracle/rlcf_oracle.py) to outputs of the reference itself (tests/golden/*.npz, produced by
oracle/make_golden.py running /root/reference/TPT in the build container).  Runs without a GPU."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import rlcf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POLICY_SEED, REWARD_SEED, VIEW_SEED, TOKEN_SEED = 0, 1, 11, 7
FAST = ["tiny_rlcf_1step", "tiny_rlcf_3step_amplify", "tiny_rlcf_process_batch", "b32_cfg1_shape",
        "tiny_rlcf_multi_reward", "tiny_rlcf_multi_reward_mean", "tiny_rlcf_reward_resize", "tiny_rlcf_min_entropy"]
SLOW = ["b16_l14_cfg2", "b16_l14_cfg2_2img", "b16_l14_cfg3_3step"]   # full config-2 / config-3 sizes (ViT-B/16 + ViT-L/14)
FULL = ["tiny_full_tune_2step", "b32_full_tune_3step"]                # only_norm=False (custom_clip.py:477-479)


def load_case(name):
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    z = np.load(path)
    cfg = ast.literal_eval(str(z["meta"]))
    return z, cfg


def oracle_setup(cfg):
    sd_p = O.make_clip_state_dict(cfg["policy"], POLICY_SEED)
    multi = isinstance(cfg["reward"], list)       # CLIPRewardsMultiple: a list of reward models
    if multi:
        sd_r = [O.make_clip_state_dict(a, s_) for a, s_ in zip(cfg["reward"], cfg["reward_seeds"])]
        vocab_r = O.ARCHS[cfg["reward"][0]][6]
    else:
        sd_r = O.make_clip_state_dict(cfg["reward"], cfg.get("reward_seed", REWARD_SEED))
        vocab_r = O.ARCHS[cfg["reward"]][6]
    tok_p = O.make_tokens(cfg["C"], O.ARCHS[cfg["policy"]][6], seed=TOKEN_SEED)
    tok_r = O.make_tokens(cfg["C"], vocab_r, seed=TOKEN_SEED)
    views = O.make_views(cfg["n_img"], cfg["V"], O.ARCHS[cfg["policy"]][1], cfg.get("view_seed", VIEW_SEED))
    ocfg = O.OracleConfig(n_views=cfg["V"], selection_p=cfg["rho"], tta_steps=cfg["steps"], sample_k=cfg["K"],
                          lr=cfg["lr"], reward_process=bool(cfg.get("reward_process", 1)),
                          process_batch=bool(cfg.get("process_batch", 0)),
                          reward_amplify=bool(cfg.get("reward_amplify", 0)),
                          reward_weights=tuple(O.ensemble_weights(cfg["confidences"])) if multi else (),
                          weighted_scores=bool(cfg.get("weighted_scores", 1)),
                          min_entropy_w=float(cfg.get("min_entropy_w", 0.0)))
    return sd_p, sd_r, tok_p, tok_r, views, ocfg


@pytest.mark.parametrize("name", FAST + SLOW + FULL)
def test_oracle_matches_reference(name):
    z, cfg = load_case(name)
    tune = "ln" if cfg.get("only_norm", True) else "full"
    torch.set_num_threads(os.cpu_count() or 1)
    sd_p, sd_r, tok_p, tok_r, views, ocfg = oracle_setup(cfg)
    cf = O.class_features(sd_p, tok_p)
    assert np.abs(cf.numpy() - z["class_feat"]).max() < 1e-5
    if isinstance(sd_r, list):
        rc = [O.class_features(r, tok_r) for r in sd_r]
        for i, f in enumerate(rc):
            assert np.abs(f.numpy() - z[f"reward_cls{i}"]).max() < 1e-5
        assert list(ocfg.reward_weights) == z["reward_weights"].tolist()
    else:
        rc = O.class_features(sd_r, tok_r)
        assert np.abs(rc.numpy() - z["reward_cls"]).max() < 1e-5
    V = cfg["V"]
    for i in range(cfg["n_img"]):
        out = O.adapt_one_image(sd_p, cf, views[i * V:(i + 1) * V], ocfg, sd_r, rc, tune=tune)
        scale = np.abs(z[f"img{i}.logits_all"]).max()
        assert np.abs(out["logits_all"].numpy() - z[f"img{i}.logits_all"]).max() < 1e-4 * scale
        if tune == "full":
            # the oracle's tune="full" mode against the reference's CLIPCLS_TTA(only_norm=False): every stored tensor
            for k in z.files:
                if k.startswith(f"img{i}.param."):
                    key = k[len(f"img{i}.param."):]
                    dd = np.abs(out["param_dict"][key].numpy() - z[k])
                    assert dd.max() <= 2.02 * cfg["lr"] * cfg["steps"], k
                    # entries whose gradient is clearly non-zero in every step (not e.g. the key third of in_proj_bias:
                    # softmax is invariant to a common shift of the keys, its gradient is rounding noise and AdamW's
                    # sign-like first steps turn that noise into +-lr)
                    g = torch.stack([gd[key].abs() / gd[key].abs().max().clamp_min(1e-30) for gd in out["grad_dicts"]])
                    well = (g.min(0).values > 1e-3).numpy()
                    assert well.any(), k
                    assert (dd[well] < 0.02 * cfg["lr"]).mean() > 0.98, k
            assert np.abs(out["logits_final"].numpy() - z[f"img{i}.logits_final"]).max() < 1e-4 * scale
            assert np.array_equal(out["selected_idx"].numpy(), z[f"img{i}.selected_idx"])
            assert np.array_equal(torch.stack(out["topk_idx"]).numpy(), z[f"img{i}.topk_idx"])
            continue
        assert np.array_equal(out["selected_idx"].numpy(), z[f"img{i}.selected_idx"])
        assert np.array_equal(torch.stack(out["topk_idx"]).numpy(), z[f"img{i}.topk_idx"])
        assert np.abs(torch.stack(out["scores"]).numpy() - z[f"img{i}.scores"]).max() < 1e-5
        rw = z[f"img{i}.rewards"]
        assert np.abs(torch.stack(out["rewards"]).numpy() - rw).max() < 1e-4 * max(1.0, np.abs(rw).max())
        assert np.abs(out["logits_final"].numpy() - z[f"img{i}.logits_final"]).max() < 1e-4 * scale
        # AdamW from an empty state moves each parameter by lr*g/(|g|+1e-8): where |g| is far above eps the oracle
        # must land within 2% of a step of the reference; where |g| ~ eps the step itself is ill-conditioned in g
        d = np.abs(out["params"].numpy() - z[f"img{i}.params"])
        gmin = torch.stack(out["grads"]).abs().min(0).values.numpy()
        assert d.max() <= 2.02 * cfg["lr"] * cfg["steps"]
        assert (gmin > 1e-5).any(), "degenerate case: all gradients vanish"
        assert d[gmin > 1e-5].max() < 0.02 * cfg["lr"]
        assert (d < 0.02 * cfg["lr"]).mean() > 0.99
        assert np.argmax(out["logits_final"].numpy()) == np.argmax(z[f"img{i}.logits_final"])


def test_weight_generator_is_deterministic():
    a = O.make_clip_state_dict("tiny-A", 3)
    b = O.make_clip_state_dict("tiny-A", 3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert O.flat_ln_params(a).numel() == (4 * 2 + 4) * 128
    assert O.flat_ln_params(O.make_clip_state_dict("ViT-B/16", 0)).numel() == 39936


def test_selection_edge_cases():
    logits = torch.randn(8, 5)
    out, idx, ent = O.select_confident_samples(logits, 0.5)
    assert out.shape == (4, 5) and torch.equal(out, logits[idx])
    assert torch.all(ent[idx][1:] >= ent[idx][:-1])
    # int(8 * 0.1) == 0 views: the reference selects nothing (SURVEY.md 8(d) config 1 note)
    out, idx, _ = O.select_confident_samples(logits, 0.1)
    assert out.shape[0] == 0
    # a single sampled class is returned unprocessed (clip_reward.py:157)
    s = torch.tensor([[0.3], [0.7]])
    assert torch.equal(O.rewards_post_process(s), s.flatten())


def prompt_layout_of(cfg, tokens, n_ctx):
    """(src_map | None, position, split_idx, learned_cls) of a prompt fixture (PromptLearner.__init__, custom_clip.py:90-99)."""
    position, split_idx = cfg.get("ctx_position", "end"), None
    words = cfg["ctx_init"].replace("_", " ").split(" ")
    if "[CLS]" in words:
        split_idx, position = words.index("[CLS]"), "middle"
    learned = bool(cfg.get("learned_cls", False))
    if position == "end" and not learned:
        return None, position, split_idx, learned
    return O.prompt_source_map(tokens, n_ctx, position, split_idx, learned), position, split_idx, learned


@pytest.mark.parametrize("name", ["tiny_prompt_rlcf_2step", "b32_prompt_rlcf", "b32_cfg1_exact", "tiny_prompt_middle",
                                  "tiny_prompt_front", "tiny_prompt_cls_word", "tiny_prompt_learned_cls"])
def test_prompt_oracle_matches_reference(name):
    """Prompt tuning (tpt_cls_rl.py / tpt_cls.py + ClipTestTimeTuning) -- oracle vs the reference's own outputs.
    b32_cfg1_exact is BASELINE.json configs[0] exactly (TPT entropy loss, TPT/tpt_cls.py:49-78)."""
    z, cfg = load_case(name)
    torch.set_num_threads(os.cpu_count() or 1)
    tpt = cfg.get("loss") == "tpt"
    sd_p = O.make_clip_state_dict(cfg["policy"], cfg.get("policy_seed", POLICY_SEED))
    tokens = torch.tensor(z["tokens"])
    ctx_init = torch.tensor(z["ctx_init"])
    assert torch.equal(ctx_init, sd_p["token_embedding.weight"][tokens[0, 1:1 + ctx_init.shape[0]]])
    sd_r = rc = None
    if not tpt:
        sd_r = O.make_clip_state_dict(cfg["reward"], cfg.get("reward_seed", REWARD_SEED))
        rc = O.class_features(sd_r, tokens)
        assert np.abs(rc.numpy() - z["reward_cls"]).max() < 1e-5
    views = O.make_views(cfg["n_img"], cfg["V"], O.ARCHS[cfg["policy"]][1], cfg.get("view_seed", VIEW_SEED))
    ocfg = O.OracleConfig(n_views=cfg["V"], selection_p=cfg["rho"], tta_steps=cfg["steps"], sample_k=cfg["K"],
                          lr=cfg["lr"], loss="tpt" if tpt else "rlcf")
    src_map, _, _, learned = prompt_layout_of(cfg, tokens, ctx_init.shape[0])
    cls_init = torch.tensor(z["cls_init"]) if learned else None
    if "prompts0" in z.files:       # the reference's assembled prompt embeddings: the layout itself, bit for bit
        emb = sd_p["token_embedding.weight"][tokens]
        if src_map is None:
            mine = torch.cat([emb[:, :1], ctx_init.expand(tokens.shape[0], -1, -1), emb[:, 1 + ctx_init.shape[0]:]], 1)
        else:
            vecs = ctx_init if cls_init is None else torch.cat([ctx_init, cls_init.reshape(tokens.shape[0], -1)])
            frozen = torch.gather(emb, 1, src_map.clamp_min(0).unsqueeze(-1).expand(-1, -1, emb.shape[-1]))
            mine = torch.where((src_map < 0).unsqueeze(-1), vecs[(-1 - src_map).clamp_min(0)], frozen)
        assert torch.equal(mine, torch.tensor(z["prompts0"]))
    V = cfg["V"]
    for i in range(cfg["n_img"]):
        out = O.adapt_one_image_prompt(sd_p, tokens, ctx_init, views[i * V:(i + 1) * V], ocfg, sd_r, rc,
                                       src_map=src_map, cls_init=cls_init)
        scale = np.abs(z[f"img{i}.logits_all"]).max()
        assert np.abs(out["logits_all"].numpy() - z[f"img{i}.logits_all"]).max() < 1e-4 * scale
        assert np.array_equal(out["selected_idx"].numpy(), z[f"img{i}.selected_idx"])
        if not tpt:
            assert np.array_equal(torch.stack(out["topk_idx"]).numpy(), z[f"img{i}.topk_idx"])
            assert np.abs(torch.stack(out["rewards"]).numpy() - z[f"img{i}.rewards"]).max() < 1e-5
        # two fp32 CPU implementations of the same arithmetic in a different operation order: 1e-4 of the logit scale,
        # plus 2e-4 of what adaptation changed (lr 5e-3 on context entries of magnitude 0.02 moves the logits by more
        # than their own scale, and AdamW's sign-like first steps amplify the last-bit differences of the gradient)
        delta = np.abs(z[f"img{i}.logits_final"][0] - z[f"img{i}.logits_all"][0]).max()
        assert np.abs(out["logits_final"].numpy() - z[f"img{i}.logits_final"]).max() < 1e-4 * scale + 2e-4 * delta
        d = np.abs(out["params"].numpy() - z[f"img{i}.params"])
        assert d.max() <= 2.02 * cfg["lr"] * cfg["steps"] and (d < 0.02 * cfg["lr"]).mean() > 0.99
