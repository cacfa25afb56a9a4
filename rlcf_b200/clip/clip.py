"""`clip.load` / `clip.tokenize` / `clip.available_models` with the TPT signatures (TPT/clip/clip.py:94-233).

load() returns the 3-tuple (model, embed_dim, preprocess) of the TPT copy (clip.py:142).  `name` may be
  * a path to an OpenAI TorchScript archive or a plain state_dict file (clip.py:121-137),
  * a key of available_models() whose file already sits in `download_root` (there is no network here, so nothing is
    ever downloaded -- a missing file raises RuntimeError like an unknown name does),
  * "synthetic:<arch>[:<seed>]" -- seeded random weights of that architecture (offline tests / benchmarks).
"""
from __future__ import annotations

import os
import warnings
from typing import List, Union

import torch

from .. import synthetic
from .model import build_model
from .simple_tokenizer import SimpleTokenizer as _Tokenizer

__all__ = ["available_models", "load", "tokenize"]

_MODELS = {  # file names of the OpenAI release (clip.py:30-40); only ViT towers are supported by the kernels
    "ViT-B/32": "ViT-B-32.pt",
    "ViT-B/16": "ViT-B-16.pt",
    "ViT-L/14": "ViT-L-14.pt",
    "ViT-L/14@336px": "ViT-L-14-336px.pt",
}
_tokenizer = None
_synthetic_loaded = False   # set by load("synthetic:..."): only then may tokenize() use the byte-level fallback


def _get_tokenizer():
    """The BPE tokenizer (RuntimeError when OpenAI's merge table is missing, unless only synthetic weights are in use)."""
    global _tokenizer
    if _tokenizer is None:
        _tokenizer = _Tokenizer(allow_byte_fallback=_synthetic_loaded)
    return _tokenizer


def available_models() -> List[str]:
    return list(_MODELS.keys())


def _transform(n_px):
    """Resize(bicubic) -> CenterCrop -> RGB -> ToTensor -> Normalize with CLIP's statistics (clip.py:79-86)."""
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
    return Compose([Resize(n_px, interpolation=InterpolationMode.BICUBIC), CenterCrop(n_px),
                    lambda image: image.convert("RGB"), ToTensor(),
                    Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])


def load(name: str, device: Union[str, torch.device] = "cuda" if torch.cuda.is_available() else "cpu",
         jit: bool = False, download_root: str = None):
    if jit:
        warnings.warn("jit=True is not supported by rlcf_b200 (the forward runs on its own CUDA kernels); ignoring")
    if name.startswith("synthetic:"):
        global _synthetic_loaded
        _synthetic_loaded = True
        parts = name.split(":")
        arch, seed = parts[1], int(parts[2]) if len(parts) > 2 else 0
        if arch not in synthetic.ARCHS:
            raise RuntimeError(f"Model {name} not found; synthetic architectures = {list(synthetic.ARCHS)}")
        state_dict = synthetic.make_state_dict(arch, seed)
    else:
        if name in _MODELS:
            model_path = os.path.join(download_root or os.path.expanduser("~/.cache/clip"), _MODELS[name])
            if not os.path.isfile(model_path):
                raise RuntimeError(f"Model {name}: {model_path} not found and this build cannot download "
                                   f"(no network); place the OpenAI checkpoint there")
        elif os.path.isfile(name):
            model_path = name
        else:
            raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
        try:
            state_dict = torch.jit.load(model_path, map_location="cpu").eval().state_dict()
        except RuntimeError:
            state_dict = torch.load(model_path, map_location="cpu")
            if "state_dict" in state_dict:
                state_dict = state_dict["state_dict"]
    embed_dim = state_dict["text_projection"].shape[1]
    model = build_model(state_dict).to(device).float()
    return model, embed_dim, _transform(model.visual.input_resolution)


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False) -> torch.LongTensor:
    """[SOT] + BPE(text) + [EOT], zero padded to context_length (clip.py:197-233)."""
    if isinstance(texts, str):
        texts = [texts]
    tok = _get_tokenizer()
    sot, eot = tok.encoder["<|startoftext|>"], tok.encoder["<|endoftext|>"]
    result = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        tokens = [sot] + tok.encode(text) + [eot]
        if len(tokens) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {texts[i]} is too long for context length {context_length}")
            tokens = tokens[:context_length]
            tokens[-1] = eot
        result[i, :len(tokens)] = torch.tensor(tokens)
    return result
