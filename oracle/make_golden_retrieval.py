"""Generates tests/golden/ret_*.npz by running the reference's retrieval TTA (/root/reference/retrieval) on
deterministic inputs.  Build container only:   python oracle/make_golden_retrieval.py [case ...]

Reference code executed unmodified: clip_ret_policy.tune_image / tune_text, custom_models.CLIPRet_TTA (forward,
parameters, momentum_update_model, reset_initial), clip_reward.get_reward_model / CLIPRewards,
lavis.models.clip_models.model.build_model_from_openai_state_dict (and with it the CLIP towers and the fp16 rounding
at load), lavis.tasks.retrieval.RetrievalTask._report_metrics.

Substituted, none of it arithmetic: the LAVIS framework around them (registry/config/runner/dataset packages need
omegaconf, iopath, timm, ... which are not installed) is replaced by empty stub modules so that the four files above
import; `load_openai_model` (reads a checkpoint file) hands seeded synthetic weights to the reference's own
build_model_from_openai_state_dict; captions are seeded synthetic token ids passed through `tokenized_prompts`
/ `tokenized_cap`; the DOWNLOAD_ROOT existence check; the `ftfy` module; registry.get_path("output_dir") for the
metrics log is pointed at a temporary directory.
"""
from __future__ import annotations

import argparse
import copy
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import rlcf_oracle as O  # noqa: E402

REF = "/root/reference/retrieval"

CASES = {
    # image -> text: the whole image encoder is tuned per query image against a gallery of captions
    "ret_i2t_tiny_3step": dict(task="image2text", policy="tiny-A", reward="tiny-B", n_query=3, n_gallery=24, K=5,
                               steps=3, lr=1e-4, momentum_update=1, update_freq=2, update_w=0.5, momentum=0.9),
    "ret_i2t_tiny_recipe": dict(task="image2text", policy="tiny-A", reward="tiny-B", n_query=2, n_gallery=40, K=20,
                                steps=8, lr=1e-6),
    # text -> image: the text tower (+ token embedding, logit_scale) is tuned per caption against a gallery of images
    "ret_t2i_tiny_3step": dict(task="text2image", policy="tiny-A", reward="tiny-B", n_query=3, n_gallery=16, K=4,
                               steps=3, lr=1e-4, momentum_update=1, update_freq=2, update_w=0.5, momentum=0.9),
    "ret_t2i_tiny_recipe": dict(task="text2image", policy="tiny-A", reward="tiny-B", n_query=2, n_gallery=30, K=12,
                                steps=8, lr=1e-6),
}
POLICY_SEED, REWARD_SEED, IMAGE_SEED, TOKEN_SEED = 0, 1, 21, 9
PARAM_STRIDE = 16     # adapted parameters are stored as every 16th element of the concatenated trainable vector


def _stub(name, path=None, **attrs):
    m = types.ModuleType(name)
    if path is not None:
        m.__path__ = [path]
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


def import_reference():
    lv = REF + "/lavis"
    sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda s: s))

    class BaseModel(torch.nn.Module):
        @property
        def device(self):
            return next(self.parameters()).device

    ident = lambda *a, **k: None  # noqa: E731
    _stub("lavis", lv)
    _stub("lavis.common", lv + "/common")
    _stub("lavis.common.utils", get_abs_path=lambda p: p, now=lambda: "golden")
    _stub("lavis.common.dist_utils", is_main_process=lambda: True, get_rank=lambda: 0, init_distributed_mode=ident)
    _stub("lavis.common.config", Config=object)
    _stub("lavis.common.logger", setup_logger=ident, MetricLogger=object)
    _stub("lavis.models", lv + "/models", BaseModel=BaseModel)
    _stub("lavis.models.base_model", BaseModel=BaseModel)
    BaseTask = type("BaseTask", (), {})
    _stub("lavis.tasks", lv + "/tasks", BaseTask=BaseTask)
    _stub("lavis.tasks.base_task", BaseTask=BaseTask)
    _stub("lavis.tasks.multimodal_classification", MultimodalClassificationTask=object)
    _stub("lavis.datasets", lv + "/datasets")
    _stub("lavis.datasets.builders")
    _stub("lavis.processors")
    _stub("lavis.runners", lv + "/runners")
    _stub("lavis.runners.runner_base", RunnerBase=object)
    _stub("lavis_evaluate", setup_seeds=ident)
    real_exists = os.path.exists
    os.path.exists = lambda p: True if p == "/YOUR/PATH" else real_exists(p)
    sys.path.insert(0, REF)
    argv, sys.argv = sys.argv, ["x"]
    try:
        import lavis.common.registry as registry            # real
        import lavis.models.clip_models.model as clip_model  # real
        import lavis.tasks.retrieval as ret_task             # real
        import custom_models, clip_reward, clip_ret_policy   # real  # noqa: E401
    finally:
        os.path.exists = real_exists
        sys.argv = argv
    return registry.registry, clip_model, ret_task, custom_models, clip_reward, clip_ret_policy


def run_case(name: str, cfg: dict, mods) -> dict:
    registry, clip_model, ret_task, custom_models, clip_reward, policy = mods
    i2t = cfg["task"] == "image2text"
    sd_p = O.make_clip_state_dict(cfg["policy"], POLICY_SEED)
    sd_r = O.make_clip_state_dict(cfg["reward"], REWARD_SEED)
    res, vocab = O.ARCHS[cfg["policy"]][1], O.ARCHS[cfg["policy"]][6]
    nq, ng = cfg["n_query"], cfg["n_gallery"]
    n_img, n_txt = (nq, ng) if i2t else (ng, nq)
    images = O.make_views(n_img, 1, res, IMAGE_SEED)              # one view per image
    tokens = O.make_tokens(n_txt, vocab, seed=TOKEN_SEED)

    def fake_loader(sd):
        def load(path, device="cpu", jit=False):
            m = clip_model.build_model_from_openai_state_dict({k: v.clone() for k, v in sd.items()}).to(device)
            return m.float()                                       # load_openai_model: `if str(device) == "cpu"`
        return load

    custom_models.load_openai_model = fake_loader(sd_p)
    clip_reward.load_openai_model = fake_loader(sd_r)
    args = argparse.Namespace(
        tta_steps=cfg["steps"], multiple_reward_models=0, reward_arch=cfg["reward"], reward_amplify=0,
        sample_k=cfg["K"], reward_process=1, process_batch=0, weighted_scores=1)
    model = custom_models.CLIPRet_TTA("cpu", arch=cfg["policy"], only_visual=i2t,
                                      momentum_update=bool(cfg.get("momentum_update", 0)),
                                      update_freq=cfg.get("update_freq", 256), update_w=cfg.get("update_w", 1.0),
                                      momentum=cfg.get("momentum", 0.9999))
    reward_model = clip_reward.get_reward_model("cpu", args)
    optimizer = torch.optim.AdamW(model.parameters(), lr=cfg["lr"], eps=1e-06, weight_decay=5e-4)   # :235
    optim_state = copy.deepcopy(optimizer.state_dict())
    scaler = torch.cuda.amp.GradScaler(init_scale=1000)

    # gallery features exactly as test_time_tune does (clip_ret_policy.py:150-160), captions pre-tokenised
    model.eval()
    with torch.no_grad():
        if i2t:
            model.set_text_features(text_features=model.get_text_features(tokenized_prompts=tokens))
            reward_model.set_text_features(tokenized_cap=tokens)
        else:
            model.set_image_features(image_features=model.get_image_features(images))
            reward_model.set_image_features(images=images)

    rec = {}
    orig_score, orig_post = reward_model.CLIPScore, reward_model.rewards_post_process

    def score(**kw):
        s = orig_score(**kw)
        idx = kw.get("text_index") if kw.get("text_index") is not None else kw.get("images_index")
        rec.setdefault("topk_idx", []).append(idx.clone())
        rec.setdefault("scores", []).append(s.clone())
        return s

    def post(cs):
        r = orig_post(cs)
        rec.setdefault("rewards", []).append(r.clone())
        return r

    reward_model.CLIPScore, reward_model.rewards_post_process = score, post
    if not i2t:
        # tune_text hands the raw caption to the tokenizer; feed the synthetic token ids of this query instead
        cur = {}
        custom_models.tokenize = lambda text: cur["tok"]
        clip_reward.tokenize = lambda text: cur["tok"]

    out = {"gallery_policy": (model.text_features if i2t else model.image_features).numpy(),
           "gallery_reward": (reward_model.text_features if i2t else reward_model.image_features).numpy()}
    rows = []
    for q in range(nq):
        rec.clear()
        if i2t:
            image = images[q:q + 1]
            policy.tune_image(image, model, reward_model, optimizer, scaler, args=args)
            out[f"q{q}.reward_query"] = reward_model.image_features.numpy().copy()
            model.eval()
            with torch.no_grad():
                logits, _ = model(image)
        else:
            cur["tok"] = tokens[q:q + 1]
            policy.tune_text("caption", model, reward_model, optimizer, scaler, args=args)
            out[f"q{q}.reward_query"] = reward_model.text_features.numpy().copy()
            model.eval()
            with torch.no_grad():
                _, logits = model(images=None, text="caption")
        rows.append(logits[0].clone())
        out[f"q{q}.score_row"] = logits[0].numpy().copy()
        out[f"q{q}.topk_idx"] = torch.stack(rec["topk_idx"]).numpy()
        out[f"q{q}.scores"] = torch.stack(rec["scores"]).numpy()
        out[f"q{q}.rewards"] = torch.stack(rec["rewards"]).numpy()
        sd_now = model.clip_model.state_dict()
        names = O.retrieval_trainable(sd_now, cfg["task"])
        out[f"q{q}.params"] = torch.cat([sd_now[n].flatten() for n in names])[::PARAM_STRIDE].numpy().copy()
        model.momentum_update_model()                              # clip_ret_policy.py:170-173
        model.reset_initial()
        optimizer.load_state_dict(optim_state)
        out[f"q{q}.initial_after"] = torch.cat([model.initial_state_dict[n].flatten() for n in names])[
            ::PARAM_STRIDE].numpy().copy()
    scores = torch.stack(rows).numpy()
    out["score_matrix"] = scores
    # recall metrics on this matrix with a synthetic ground truth (two captions per image)
    n_i, n_t = (nq, ng) if i2t else (ng, nq)
    rng = np.random.RandomState(5)
    if i2t:
        img2txt = [sorted(rng.choice(n_t, 2, replace=False).tolist()) for _ in range(n_i)]
        s_t2i = rng.randn(n_t, n_i).astype(np.float32)
        txt2img = rng.randint(0, n_i, n_t).tolist()
        s_i2t = scores
    else:
        txt2img = rng.randint(0, n_i, n_t).tolist()
        s_i2t = rng.randn(n_i, n_t).astype(np.float32)
        img2txt = [sorted(rng.choice(n_t, min(2, n_t), replace=False).tolist()) for _ in range(n_i)]
        s_t2i = scores
    with tempfile.TemporaryDirectory(dir=HERE) as tmp:
        registry.mapping["paths"]["output_dir"] = tmp
        metrics = ret_task.RetrievalTask._report_metrics(s_i2t, s_t2i, txt2img, img2txt)
    out["metrics_keys"] = np.array(sorted(metrics))
    out["metrics_vals"] = np.array([metrics[k] for k in sorted(metrics)], dtype=np.float64)
    out["metrics_s_i2t"], out["metrics_s_t2i"] = s_i2t, s_t2i
    out["metrics_txt2img"] = np.array(txt2img)
    out["metrics_img2txt"] = np.array(img2txt)
    out["param_stride"] = np.array(PARAM_STRIDE)
    out["cfg_keys"] = np.array(sorted(cfg))
    out["cfg_vals"] = np.array([str(cfg[k]) for k in sorted(cfg)])
    return out


def main():
    names = sys.argv[1:] or list(CASES)
    mods = import_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for n in names:
        torch.manual_seed(0)
        out = run_case(n, CASES[n], mods)
        path = os.path.join(ROOT, "tests", "golden", n + ".npz")
        np.savez_compressed(path, **out)
        print(n, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", flush=True)


if __name__ == "__main__":
    main()
