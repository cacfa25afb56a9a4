"""Reward model with the reference's surface (TPT/clip_reward.py): a frozen CLIP that scores sampled predictions.

get_reward_model(device, args) -> CLIPRewards (clip_reward.py:29-40) or, with --multiple_reward_models 1, the ensemble
CLIPRewardsMultiple (clip_reward.py:180-307).  The ensemble works for any list of ViT CLIPs (views are resized to a
member's own resolution, clip_reward.py:133-134 / 262-264); the reference's hard-coded list ["ViT-L/14@336px",
"RN50x64", "ViT-L/14"] contains a ModifiedResNet, which is out of scope -- pass the
architectures as a comma-separated --reward_arch instead.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import clip
from ._lib import RlcfError

DOWNLOAD_ROOT = None

CONFIDECES = {"ViT-L/14@336px": 10, "ViT-L/14": 5, "RN50x64": 3, "ViT-B/16": 1}


def get_reward_model(device, args):
    if getattr(args, "multiple_reward_models", 0):
        archs = [a.strip() for a in str(args.reward_arch).split(",") if a.strip()]
        if len(archs) < 2:
            raise NotImplementedError(
                "--multiple_reward_models 1: the reference's ensemble [ViT-L/14@336px, RN50x64, ViT-L/14] "
                "(clip_reward.py:31) contains a ResNet tower; give 2-4 ViT architectures as "
                "--reward_arch 'ViT-L/14,ViT-B/16'")
        return CLIPRewardsMultiple(device, arch=archs, classification=True, amplify_rewards=args.reward_amplify,
                                   sample_k=args.sample_k, reward_process=args.reward_process,
                                   process_batch=args.process_batch, weighted_scores=getattr(args, "weighted_scores", 1))
    return CLIPRewards(device, arch=args.reward_arch, classification=True, amplify_rewards=args.reward_amplify,
                       sample_k=args.sample_k, reward_process=args.reward_process,
                       process_batch=args.process_batch)


get_clip_reward = get_reward_model  # the name BASELINE.json's north_star uses


class BaseRewards(nn.Module):
    @torch.no_grad()
    def set_class_features(self, classnames=None, tokenized_classes=None):
        self.class_features = self.extract_text_features(captions=classnames, tokenized_cap=tokenized_classes)

    @torch.no_grad()
    def set_image_features(self, images):
        self.image_features = self.extract_image_features(images)

    @torch.no_grad()
    def confidence_gap(self, predictions):
        value, _ = torch.topk(predictions, 2, dim=-1)
        gap = value[:, 0] - value[:, 1]
        return gap - torch.mean(gap)


def engine_inputs(reward_model):
    """(reward tower(s), class features, weights) in the form the engines take: single objects for CLIPRewards, lists
    (+ one weight per model) for CLIPRewardsMultiple."""
    if isinstance(reward_model, CLIPRewardsMultiple):
        n = reward_model.n_model
        w = reward_model.weights if reward_model.weighted_scores else [1.0 / n] * n
        return [m.visual.tower() for m in reward_model.clip_models], reward_model.class_features, tuple(w)
    return reward_model.clip_model.visual.tower(), reward_model.class_features, ()


class CLIPRewards(BaseRewards):
    def __init__(self, device, arch="ViT-B/16", clipscore_weight=2.5, classification=True, amplify_rewards=False,
                 sample_k=5, reward_process=True, process_batch=False, default_resolutions=224) -> None:
        super().__init__()
        self.default_resolutions = default_resolutions
        self.clip_model, self.embed_dim, self.preprocess = clip.load(arch, device=device, download_root=DOWNLOAD_ROOT)
        self.resolutions = self.clip_model.visual.input_resolution
        self.clipscore_weight = clipscore_weight
        self.device = device
        self.classification = classification
        self.class_features = None
        self.image_features = None
        self.amplify_rewards = amplify_rewards
        self.sample_k = sample_k
        self.reward_process = reward_process
        self.process_batch = process_batch
        self.clip_model.eval()

    @torch.no_grad()
    def CLIPScore(self, class_index, images=None, image_features=None, captions=None, tokenized_cap=None,
                  text_features=None, pairwise=True):
        """max(0, w * cos(image, class text)) for the sampled class indices (clip_reward.py:111-128)."""
        text_features = self.class_features[class_index]
        image_features = torch.repeat_interleave(self.image_features, self.sample_k, dim=0)
        if pairwise:
            similarity = self.clipscore_weight * text_features @ image_features.t()
        else:
            similarity = self.clipscore_weight * torch.sum(text_features * image_features, dim=-1)
        return torch.maximum(similarity, torch.zeros_like(similarity)).squeeze()

    @torch.no_grad()
    def extract_image_features(self, images):
        if self.resolutions != self.default_resolutions:
            images = nn.functional.interpolate(images, size=self.resolutions, mode="bicubic", align_corners=True)
        image_features = self.clip_model.encode_image(images).float()
        return image_features / image_features.norm(dim=1, keepdim=True)

    @torch.no_grad()
    def extract_text_features(self, captions=None, tokenized_cap=None):
        if captions is not None:
            caption_tokens = clip.tokenize(captions, truncate=True).to(self.device)
            text_features = self.clip_model.encode_text(caption_tokens).float()
        if tokenized_cap is not None:
            text_features = self.clip_model.encode_text(tokenized_cap).float()
        if captions is None and tokenized_cap is None:
            raise RlcfError("extract_text_features needs captions or tokenized_cap")
        return text_features / text_features.norm(dim=1, keepdim=True)

    @torch.no_grad()
    def rewards_post_process(self, clip_score):
        """Baseline subtraction (and optional standardisation) over the last dim (clip_reward.py:152-165)."""
        if clip_score.shape[-1] > 1 and self.reward_process:
            mean = torch.mean(clip_score, dim=-1, keepdim=True)
            std = torch.std(clip_score, dim=-1, keepdim=True) + 1e-5 if self.amplify_rewards else 1.0
            clip_score = (clip_score - mean) / std
        return clip_score.flatten()

    @torch.no_grad()
    def calulate_similarity(self):  # (sic) reference spelling, clip_reward.py:167
        logit_scale = self.clip_model.logit_scale.exp()
        logits_per_image = logit_scale * self.image_features @ self.class_features.t()
        return logits_per_image, logits_per_image.t()


def _confidence(arch: str) -> float:
    key = arch.split(":")[1] if arch.startswith("synthetic:") else arch
    return float(CONFIDECES.get(key, 1))


class CLIPRewardsMultiple(BaseRewards):
    """Ensemble of frozen CLIPs (clip_reward.py:180-307): a sample's CLIPScore is the confidence-weighted sum (or the
    mean) of the members' scores.  Members must be ViT models (views are resized to each member's resolution)."""

    def __init__(self, device, arch=("ViT-B/16", "ViT-L/14"), clipscore_weight=2.5, classification=True,
                 amplify_rewards=False, sample_k=5, reward_process=True, process_batch=True, weighted_scores=True,
                 default_resolutions=224) -> None:
        super().__init__()
        if not 2 <= len(arch) <= 4:
            raise RlcfError("CLIPRewardsMultiple takes 2..4 reward models")
        models, weights = [], []
        self.preprocess, self.resolutions = [], []
        self.default_resolutions = default_resolutions
        for ar in arch:
            clip_model, _, preprocess = clip.load(ar, device=device, download_root=DOWNLOAD_ROOT)
            models.append(clip_model)
            self.preprocess.append(preprocess)
            self.resolutions.append(clip_model.visual.input_resolution)
            weights.append(_confidence(ar))
        self.clip_models = nn.ModuleList(models)
        self.n_model = len(models)
        self.weights = [round(x / sum(weights), 2) for x in weights]       # clip_reward.py:207
        self.clipscore_weight = clipscore_weight
        self.device = device
        self.classification = classification
        self.class_features = None
        self.image_features = None
        self.amplify_rewards, self.sample_k = amplify_rewards, sample_k
        self.reward_process, self.process_batch, self.weighted_scores = reward_process, process_batch, weighted_scores
        self.clip_models.eval()

    @torch.no_grad()
    def CLIPScore(self, class_index, images=None, image_features=None, captions=None, tokenized_cap=None,
                  text_features=None, pairwise=False):
        if pairwise:
            raise NotImplementedError    # as the reference (clip_reward.py:237-238)
        all_scores = []
        for i in range(self.n_model):
            t = self.class_features[i][class_index]
            f = torch.repeat_interleave(self.image_features[i], self.sample_k, dim=0)
            sim = self.clipscore_weight * torch.sum(t * f, dim=-1)
            all_scores.append(torch.maximum(sim, torch.zeros_like(sim)).squeeze())
        scores = torch.stack(all_scores, dim=0)
        if self.weighted_scores:
            w = torch.tensor(self.weights, device=scores.device, dtype=scores.dtype).unsqueeze(1)
            return torch.sum(w * scores, dim=0)
        return torch.mean(scores, dim=0)

    @torch.no_grad()
    def extract_image_features(self, images):
        feats = []
        for m, res in zip(self.clip_models, self.resolutions):
            x = images
            if res != images.shape[-1]:   # clip_reward.py:262-264 (the engines use rlcf_bicubic_resize for this)
                x = nn.functional.interpolate(images, size=res, mode="bicubic", align_corners=True)
            f = m.encode_image(x).float()
            feats.append(f / f.norm(dim=1, keepdim=True))
        return feats

    @torch.no_grad()
    def extract_text_features(self, captions=None, tokenized_cap=None):
        if captions is not None:
            tokenized_cap = clip.tokenize(captions, truncate=True).to(self.device)
        if tokenized_cap is None:
            raise RlcfError("extract_text_features needs captions or tokenized_cap")
        feats = []
        for m in self.clip_models:
            f = m.encode_text(tokenized_cap).float()
            feats.append(f / f.norm(dim=1, keepdim=True))
        return feats

    rewards_post_process = CLIPRewards.rewards_post_process
