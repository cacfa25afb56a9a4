"""Retrieval TTA with the reference's surface (retrieval/clip_ret_policy.py, retrieval/custom_models.py,
retrieval/clip_reward.py, retrieval/lavis/tasks/retrieval.py:52-107) on top of rlcf_b200.retrieval's query engines.

What the reference takes from LAVIS (dataset builders, Config, RunnerBase) is outside the hot path; the driver here
takes the two things it uses from them: a dataset object with `.text` (captions or a token tensor), `.image_tensor` or
an iterable of {"image": tensor} batches, `.txt2img`, `.img2txt`.

    model = CLIPRet_TTA(device, arch, only_visual=(task == "image2text"), momentum_update=..., ...)
    reward_model = get_reward_model(device, args)
    s_i2t, s_t2i = test_time_tune(dataset, device, model, reward_model, args=args)
    metrics = report_metrics(s_i2t, s_t2i, dataset.txt2img, dataset.img2txt)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import clip
from . import engine as E
from . import retrieval as R
from ._lib import RlcfError

DOWNLOAD_ROOT = None


def openai_rounding(sd: dict) -> dict:
    """load_openai_model (retrieval/lavis/models/clip_models/model.py:763-791,869-871) converts the model to fp16
    BEFORE load_state_dict and CLIPRet_TTA calls .float() afterwards (custom_models.py:40-41): Conv / Linear /
    MultiheadAttention weights and biases, visual.proj and text_projection take fp16-representable values.  A no-op
    for OpenAI archives (stored in fp16)."""
    out = {}
    for k, v in sd.items():
        rounded = (k.endswith(("conv1.weight", "in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias",
                               "c_fc.weight", "c_fc.bias", "c_proj.weight", "c_proj.bias"))
                   or k in ("visual.proj", "text_projection"))
        out[k] = v.half().float() if rounded else v.clone()
    return out


def _load_state(arch, device):
    name = arch if (":" in arch or "/" in arch) else arch.replace("ViT-B-", "ViT-B/").replace("ViT-L-", "ViT-L/")
    model, _, _ = clip.load(name, device=device, download_root=DOWNLOAD_ROOT)
    return openai_rounding({k: v.detach() for k, v in model.state_dict().items()})


class CLIPRet_TTA(nn.Module):
    """custom_models.py:28-160.  Holds the CLIP state, the candidate features of the other modality and the
    momentum state; tune_image / tune_text adapt through a query engine built on first use."""

    def __init__(self, device, arch="ViT-B-16", only_visual=True, momentum_update=False, update_freq=256, update_w=1.0,
                 momentum=0.9999, state_dict=None):
        super().__init__()
        self.device = torch.device(device)
        self.state = {k: v.to(self.device) for k, v in (state_dict or _load_state(arch, self.device)).items()}
        self.only_visual = only_visual
        self.momentum_update, self.update_freq, self.update_w, self.momentum = momentum_update, update_freq, update_w, momentum
        self.text_features = None
        self.image_features = None
        self._visual = E.prepare_visual(self.state)
        self._text = E.prepare_text(self.state)
        self._engine = None
        self._engine_key = None

    @property
    def logit_scale(self):
        return self.state["logit_scale"]

    # ---- features of the un-adapted model (gallery side) -------------------------------------------------------
    @torch.no_grad()
    def get_text_features(self, text=None, tokenized_prompts=None):
        if tokenized_prompts is None:
            if text is None:
                raise RlcfError("get_text_features needs text or tokenized_prompts")
            tokenized_prompts = clip.tokenize(text, truncate=True)
        # fp16 tensor-core tower on purpose: the retrieval CLIP is loaded in fp16 by the reference
        # (lavis/models/clip_models/model.py:763-791), gallery features there carry fp16 operand rounding too
        return E.text_features(self._text, tokenized_prompts.to(self.device), precise=False)

    @torch.no_grad()
    def get_image_features(self, images):
        return E.image_features(self._visual, images.to(self.device))

    def set_image_features(self, images=None, image_features=None):
        self.image_features = self.get_image_features(images) if images is not None else image_features
        self._retire_engine()

    def _retire_engine(self):
        """The gallery changed: the engine must be rebuilt, but its momentum state is carried into the next one."""
        if self._engine is not None:
            self._retired_engine = self._engine
        self._engine = None

    def set_text_features(self, text=None, tokenized_prompts=None, text_features=None):
        if text is not None or tokenized_prompts is not None:
            self.text_features = self.get_text_features(text, tokenized_prompts)
        else:
            if text_features is None:
                raise RlcfError("set_text_features needs text, tokenized_prompts or text_features")
            self.text_features = text_features
        self._retire_engine()

    @torch.no_grad()
    def forward(self, images=None, text=None, tokenized_prompts=None):
        image_features = self.get_image_features(images) if images is not None else self.image_features
        text_features = (self.get_text_features(text, tokenized_prompts)
                         if text is not None or tokenized_prompts is not None else self.text_features)
        logit_scale = self.logit_scale.exp()
        logits_per_image = logit_scale * image_features @ text_features.t()
        return logits_per_image, logits_per_image.t()

    # ---- adaptation --------------------------------------------------------------------------------------------
    def engine(self, reward_model, args, n_query):
        """The query engine for this model / reward model / hyper-parameters (rebuilt when any of them changes)."""
        n_query = 1 if self.momentum_update else n_query
        cfg = R.RetrievalConfig(tta_steps=args.tta_steps, sample_k=reward_model.sample_k, lr=args.lr,
                                weight_decay=args.weight_decay, reward_process=bool(reward_model.reward_process),
                                process_batch=bool(reward_model.process_batch),
                                reward_amplify=bool(reward_model.amplify_rewards),
                                clipscore_weight=reward_model.clipscore_weight, momentum_update=bool(self.momentum_update),
                                update_freq=self.update_freq, update_w=self.update_w, momentum=self.momentum)
        # every hyper-parameter the engine bakes in is part of the key; the reward model is pinned by the reference kept
        # next to the engine (so its id() cannot be recycled while the entry is alive)
        key = (id(reward_model), n_query, tuple(sorted(vars(cfg).items())))
        if self._engine is not None and self._engine_key == key:
            return self._engine
        if self.only_visual:
            if self.text_features is None or reward_model.text_features is None:
                raise RlcfError("image->text: set_text_features on the model and the reward model first")
            eng = R.ImageQueryEngine(self.state, self.text_features, float(self.logit_scale.exp()), cfg, n_query,
                                     reward_model.visual_tower(), reward_model.text_features)
        else:
            if self.image_features is None or reward_model.image_features is None:
                raise RlcfError("text->image: set_image_features on the model and the reward model first")
            eng = R.TextQueryEngine(self.state, self.image_features, cfg, n_query, reward_model.text_tower(),
                                    reward_model.image_features)
        # the momentum (EMA) state belongs to the MODEL in the reference (custom_models.py: momentum_state_dict,
        # update_counter live on CLIPRet_TTA) and must survive an engine rebuild
        old = self._engine if self._engine is not None else getattr(self, "_retired_engine", None)
        if old is not None and self.momentum_update and type(old) is type(eng):
            for name in ("ema_ln", "ema_rest", "ema_tab", "init_ln", "init_rest", "init_tab"):
                a, b = getattr(eng, name, None), getattr(old, name, None)
                if a is not None and b is not None and a.shape == b.shape:
                    a.copy_(b)
            eng.update_counter = getattr(old, "update_counter", 0)
        self._retired_engine = None
        self._engine, self._engine_key, self._engine_refs = eng, key, (reward_model,)
        return eng

    def momentum_update_model(self):
        if self._engine is not None:
            self._engine.momentum_update(0)

    def reset_initial(self):
        """The engines restore the initial weights at the start of every query; nothing to do between queries."""

    def parameters(self, recurse: bool = True):
        vis = self.only_visual
        return [v for k, v in self.state.items() if k.startswith("visual.") == vis]


class CLIPRewards(nn.Module):
    """retrieval/clip_reward.py:92-230: a frozen CLIP scoring (image, text) pairs."""

    def __init__(self, device, arch="ViT-L-14", clipscore_weight=2.5, amplify_rewards=False, sample_k=5,
                 reward_process=True, process_batch=False, state_dict=None):
        super().__init__()
        self.device = torch.device(device)
        self.state = {k: v.to(self.device) for k, v in (state_dict or _load_state(arch, self.device)).items()}
        self.clipscore_weight = clipscore_weight
        self.amplify_rewards, self.sample_k = amplify_rewards, sample_k
        self.reward_process, self.process_batch = reward_process, process_batch
        self.text_features = None
        self.image_features = None
        self._visual = self._text = None

    def visual_tower(self):
        if self._visual is None:
            self._visual = E.prepare_visual(self.state)
        return self._visual

    def text_tower(self):
        if self._text is None:
            self._text = E.prepare_text(self.state)
        return self._text

    @torch.no_grad()
    def extract_image_features(self, images):
        return E.image_features(self.visual_tower(), images.to(self.device))

    @torch.no_grad()
    def extract_text_features(self, captions=None, tokenized_cap=None):
        if captions is not None:
            tokenized_cap = clip.tokenize(captions, truncate=True)
        if tokenized_cap is None:
            raise RlcfError("extract_text_features needs captions or tokenized_cap")
        return E.text_features(self.text_tower(), tokenized_cap.to(self.device), precise=False)

    def set_image_features(self, images=None, image_features=None):
        self.image_features = self.extract_image_features(images) if images is not None else image_features

    def set_text_features(self, captions=None, tokenized_cap=None, text_features=None):
        if captions is not None or tokenized_cap is not None:
            self.text_features = self.extract_text_features(captions, tokenized_cap)
        else:
            self.text_features = text_features

    @torch.no_grad()
    def set_many_text_features(self, texts, text_bs=128):
        """All gallery captions (a list of strings or a token tensor [N, 77]), text_bs at a time."""
        feats = []
        for i in range(0, len(texts), text_bs):
            chunk = texts[i:i + text_bs]
            feats.append(self.extract_text_features(tokenized_cap=chunk) if torch.is_tensor(chunk)
                         else self.extract_text_features(captions=chunk))
        self.text_features = torch.cat(feats, dim=0)

    @torch.no_grad()
    def set_image_features_with_dataloder(self, data_loader):  # (sic) reference spelling, clip_reward.py:214
        self.image_features = torch.cat([self.extract_image_features(s["image"]) for s in data_loader], dim=0)

    @torch.no_grad()
    def CLIPScore(self, text_index=None, images_index=None, pairwise=True):
        text_features = (self.text_features[text_index] if text_index is not None
                         else torch.repeat_interleave(self.text_features, self.sample_k, dim=0))
        image_features = (self.image_features[images_index] if images_index is not None
                          else torch.repeat_interleave(self.image_features, self.sample_k, dim=0))
        if pairwise:
            similarity = self.clipscore_weight * text_features @ image_features.t()
        else:
            similarity = self.clipscore_weight * torch.sum(text_features * image_features, dim=-1)
        return torch.maximum(similarity, torch.zeros_like(similarity)).squeeze()

    @torch.no_grad()
    def rewards_post_process(self, clip_score):
        if clip_score.shape[-1] > 1 and self.reward_process:
            mean = torch.mean(clip_score, dim=-1, keepdim=True)
            std = torch.std(clip_score, dim=-1, keepdim=True) + 1e-5 if self.amplify_rewards else 1.0
            clip_score = (clip_score - mean) / std
        return clip_score.flatten()


def get_reward_model(device, args):
    """retrieval/clip_reward.py:32-45."""
    if getattr(args, "multiple_reward_models", 0):
        raise NotImplementedError("CLIPRewardsMultiple (RN50x64 + ViT-L/14@336px ensemble) is out of scope")
    return CLIPRewards(device, arch=args.reward_arch, amplify_rewards=args.reward_amplify, sample_k=args.sample_k,
                       reward_process=args.reward_process, process_batch=args.process_batch)


def build_parser():
    """Flags of retrieval/params.py:10-45 (names and defaults) plus --queries_per_step."""
    import argparse
    p = argparse.ArgumentParser(description="Retrieval test-time adaptation with CLIP reward, B200-native")
    p.add_argument("--retrieval_task", type=str, default="image2text", choices=["image2text", "text2image"])
    p.add_argument("--arch", type=str, default="ViT-B-16")
    p.add_argument("--output", type=str, default="exp_ret")
    p.add_argument("--lr", default=1e-6, type=float)
    p.add_argument("--weight_decay", default=5e-4, type=float)
    p.add_argument("--tta_steps", default=1, type=int)
    p.add_argument("--sample_k", type=int, default=5)
    p.add_argument("--multiple_reward_models", type=int, default=0)
    p.add_argument("--reward_arch", type=str, default="ViT-L-14")
    p.add_argument("--reward_process", type=int, default=1)
    p.add_argument("--process_batch", type=int, default=0)
    p.add_argument("--reward_amplify", type=int, default=0)
    p.add_argument("--weighted_scores", type=int, default=1)
    p.add_argument("--momentum_update", type=int, default=0)
    p.add_argument("--update_freq", type=int, default=256)
    p.add_argument("--update_w", type=float, default=1.0)
    p.add_argument("--tta_momentum", type=float, default=0.9999)
    p.add_argument("--queries_per_step", type=int, default=16, help="independent queries adapted per launch sequence")
    return p


# ------------------------------------------------------------------------------------------------ tuning functions
def tune_image(image, model, reward_model, optimizer=None, scaler=None, args=None):
    """clip_ret_policy.py:76-103 for a batch of independent query images [Q,3,H,W]; the adapted parameters stay in the
    engine (model.engine(...).export_params(q)).  `optimizer` / `scaler` are accepted for signature compatibility:
    AdamW (eps 1e-6, clip_ret_policy.py:235) and a static loss scale run inside the kernels."""
    eng = model.engine(reward_model, args, image.shape[0])
    eng.tune(image.to(model.device).float().contiguous())
    return eng


def tune_text(text, model, reward_model, optimizer=None, scaler=None, args=None):
    """clip_ret_policy.py:106-137; `text` is a caption, a list of captions or a token tensor [Q, 77]."""
    tokens = text if torch.is_tensor(text) else clip.tokenize(text, truncate=True)
    eng = model.engine(reward_model, args, tokens.shape[0])
    eng.tune(tokens.to(model.device))
    return eng


def shard_indices(n: int, rank: int, world: int):
    """Queries of rank r: r, r + world, ...."""
    return list(range(rank, n, world))


def gather_score_rows(local_rows: torch.Tensor, n_total: int, rank: int, world: int, group=None) -> torch.Tensor:
    """all_gather of the per-rank score rows into the [n_total, C] matrix of the strided sharding above.  Ranks hold
    ceil or floor(n_total / world) rows; shorter ones are padded for the collective."""
    if world == 1:
        return local_rows
    import torch.distributed as dist
    per = (n_total + world - 1) // world
    pad = torch.full((per, local_rows.shape[1]), -100.0, dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.full((n_total, local_rows.shape[1]), -100.0, dtype=local_rows.dtype, device=local_rows.device)
    for r in range(world):
        idx = shard_indices(n_total, r, world)
        out[idx] = parts[r][:len(idx)]
    return out


def test_time_tune(dataset, device, model, reward_model=None, scaler=None, optimizer=None, optim_state=None,
                   text_bs=128, args=None, queries_per_step=16, rank=0, world=1):
    """clip_ret_policy.py:140-195.  Returns (score_matrix_i2t, score_matrix_t2i) as numpy arrays; the direction that
    was not run is filled with -100 as in the reference.  dataset.text: list of captions or token tensor [Nt, 77];
    dataset.image_tensor: [Ni, 3, H, W] (normalised pixels)."""
    text = dataset.text
    tokens = text if torch.is_tensor(text) else clip.tokenize(list(text), truncate=True)
    images = dataset.image_tensor
    n_img, n_txt = images.shape[0], tokens.shape[0]
    dev = model.device
    if model.only_visual:
        model.set_text_features(tokenized_prompts=tokens)
        reward_model.set_many_text_features(tokens, text_bs=text_bs)
        mine = shard_indices(n_img, rank, world)
        rows = torch.empty(len(mine), n_txt, dtype=torch.float32, device=dev)
        q = 1 if model.momentum_update else queries_per_step
        for s in range(0, len(mine), q):
            idx = mine[s:s + q]
            batch = images[idx].to(dev).float()
            if len(idx) < q:   # last, ragged batch: pad with repeats of its first query (discarded below)
                batch = torch.cat([batch, batch[:1].expand(q - len(idx), -1, -1, -1)], dim=0)
            eng = tune_image(batch, model, reward_model, optimizer, scaler, args=args)
            rows[s:s + len(idx)] = eng.predict(batch.contiguous())[:len(idx)]
            model.momentum_update_model()
        s_i2t = gather_score_rows(rows, n_img, rank, world)
        s_t2i = torch.full((n_txt, n_img), -100.0, device=dev)
    else:
        model.set_image_features(images=images)
        reward_model.set_image_features(images=images)
        mine = shard_indices(n_txt, rank, world)
        rows = torch.empty(len(mine), n_img, dtype=torch.float32, device=dev)
        q = 1 if model.momentum_update else queries_per_step
        for s in range(0, len(mine), q):
            idx = mine[s:s + q]
            batch = tokens[idx]
            if len(idx) < q:
                batch = torch.cat([batch, batch[:1].expand(q - len(idx), -1)], dim=0)
            eng = tune_text(batch, model, reward_model, optimizer, scaler, args=args)
            rows[s:s + len(idx)] = eng.predict()[:len(idx)]
            model.momentum_update_model()
        s_t2i = gather_score_rows(rows, n_txt, rank, world)
        s_i2t = torch.full((n_img, n_txt), -100.0, device=dev)
    return s_i2t.detach().cpu().numpy(), s_t2i.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------ recall metrics
def _ranks_desc(scores: np.ndarray) -> np.ndarray:
    """rank[i, j] = position of candidate j when row i is sorted by descending score (0 = best).  Ties are ordered by
    descending index, the order np.argsort(score)[::-1] gives for a stable sort (the reference's quicksort leaves the
    order of exactly equal scores unspecified)."""
    order = np.argsort(scores, axis=1, kind="stable")[:, ::-1]
    ranks = np.empty_like(order)
    rows = np.arange(scores.shape[0])[:, None]
    ranks[rows, order] = np.arange(scores.shape[1])[None, :]
    return ranks


def report_metrics(scores_i2t, scores_t2i, txt2img, img2txt) -> dict:
    """RetrievalTask._report_metrics (retrieval/lavis/tasks/retrieval.py:52-107): recall@1/5/10 for image->text (best
    rank over the image's ground-truth captions) and text->image, their means, and agg_metrics = mean text recall."""
    r_i2t = _ranks_desc(np.asarray(scores_i2t))
    best = np.array([min(r_i2t[i, j] for j in img2txt[i]) for i in range(r_i2t.shape[0])])
    tr = [100.0 * float(np.sum(best < k)) / len(best) for k in (1, 5, 10)]
    r_t2i = _ranks_desc(np.asarray(scores_t2i))
    got = r_t2i[np.arange(r_t2i.shape[0]), np.asarray(txt2img)]
    ir = [100.0 * float(np.sum(got < k)) / len(got) for k in (1, 5, 10)]
    tr_mean, ir_mean = sum(tr) / 3, sum(ir) / 3
    return {"txt_r1": tr[0], "txt_r5": tr[1], "txt_r10": tr[2], "txt_r_mean": tr_mean, "img_r1": ir[0], "img_r5": ir[1],
            "img_r10": ir[2], "img_r_mean": ir_mean, "r_mean": (tr_mean + ir_mean) / 2, "agg_metrics": sum(tr) / 3}
