"""Host-side orchestration of the sm_100a kernels: transformer towers and the batched RLCF adaptation step.

Layout in HBM (all row-major):
  * token rows: row = sequence * L + token, residual stream fp32 [rows, d]; GEMM operands fp16
  * frozen GEMM weights fp16 [out, in] (PyTorch Linear layout = K-major "B" operand) plus, when a backward is
    needed, a transposed fp16 copy [in, out] that serves as the B operand of the dgrad GEMM
  * trainable LayerNorm slice: one flat fp32 vector per parameter set,
        [ln_pre.w, ln_pre.b, (ln_1.w, ln_1.b, ln_2.w, ln_2.b) x layers, ln_post.w, ln_post.b]
    (39 936 floats for ViT-B/16); `n_sets` copies when every test image owns its parameters.

Reference call stack being replaced: TPT/tpt_cls_rl.py:47-79 (test_time_tuning), TPT/clip/model.py:171-240.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import torch

from . import ops
from ._lib import EPI_F16, EPI_F32, EPI_GELU_BWD_F16, EPI_GELU_F16, EPI_RESID_F32, RlcfError

TRUNCATE_TEXT = os.environ.get("RLCF_TEXT_TRUNCATE", "1") != "0"   # PromptEngine: drop the padding after the last EOT
PRUNE_LAST = os.environ.get("RLCF_PRUNE_LAST", "1") != "0"   # TowerRunner: last block on the class-token rows only
N_SLOTS = 32  # gradient partial slots per parameter set = LN-backward blocks per image (deterministic reduction)


def _round_up(x, m):
    return (x + m - 1) // m * m


@dataclass
class LayerWeights:
    wqkv: torch.Tensor  # fp16 [3d, d]
    bqkv: torch.Tensor  # fp32 [3d]
    wo: torch.Tensor    # fp16 [d, d]
    bo: torch.Tensor
    wfc: torch.Tensor   # fp16 [4d, d]
    bfc: torch.Tensor
    wproj: torch.Tensor  # fp16 [d, 4d]
    bproj: torch.Tensor
    # transposed copies (dgrad B operands), only when prepared with need_grad
    wqkv_t: torch.Tensor | None = None  # [d, 3d]
    wo_t: torch.Tensor | None = None    # [d, d]
    wfc_t: torch.Tensor | None = None   # [d, 4d]
    wproj_t: torch.Tensor | None = None  # [4d, d]
    # fp32 originals [out, in] (text towers only): operands of the once-per-dataset fp32 class-feature path
    w32: dict | None = None


@dataclass
class TowerWeights:
    """Device-resident, kernel-ready weights of one transformer tower (visual or text)."""
    kind: str            # "visual" | "text"
    d: int
    heads: int
    n_layers: int
    L: int               # tokens per sequence
    E: int               # output embedding dim
    layers: list = field(default_factory=list)
    proj: torch.Tensor | None = None      # fp32 [d, E]
    ln_flat: torch.Tensor | None = None   # fp32 [P]
    has_ln_pre: bool = True
    # visual only
    patch: int = 0
    resolution: int = 0
    k_pad: int = 0
    conv_w: torch.Tensor | None = None    # fp16 [d, k_pad]
    cls: torch.Tensor | None = None       # fp32 [d]
    pos: torch.Tensor | None = None       # fp32 [L, d]
    # text only
    tok_emb: torch.Tensor | None = None   # fp32 [vocab, d]
    # per-sample weights (every test sample owns the whole tower: full tuning after its first step, retrieval TTA):
    # the GEMM weights are then 3-D [G, out, in] views, biases [G, n], and cls/pos (embed_stride) and proj
    # (proj_stride) are the first sample's tensors with the stride, in floats, to the next sample's.
    embed_stride: int = 0
    proj_stride: int = 0

    @property
    def P(self):
        return self.ln_flat.numel()

    def ln_off(self, name: str, layer: int = 0) -> int:
        """Offset (floats) of a LayerNorm's gamma inside the flat vector; beta follows at +d."""
        base = 2 * self.d if self.has_ln_pre else 0
        if name == "ln_pre":
            if not self.has_ln_pre:
                raise RlcfError("tower has no ln_pre")
            return 0
        if name == "ln_1":
            return base + layer * 4 * self.d
        if name == "ln_2":
            return base + layer * 4 * self.d + 2 * self.d
        if name in ("ln_post", "ln_final"):
            return base + self.n_layers * 4 * self.d
        raise KeyError(name)

    def ln_names(self, prefix: str):
        """(state_dict key, offset) pairs in flat order, for packing/unpacking."""
        out = []
        if self.has_ln_pre:
            out += [(f"{prefix}ln_pre.weight", 0), (f"{prefix}ln_pre.bias", self.d)]
        for l in range(self.n_layers):
            o1, o2 = self.ln_off("ln_1", l), self.ln_off("ln_2", l)
            rb = f"{prefix}transformer.resblocks.{l}."
            out += [(rb + "ln_1.weight", o1), (rb + "ln_1.bias", o1 + self.d),
                    (rb + "ln_2.weight", o2), (rb + "ln_2.bias", o2 + self.d)]
        last = "ln_post" if self.kind == "visual" else "ln_final"
        o = self.ln_off("ln_post")
        out += [(f"{prefix}{last}.weight", o), (f"{prefix}{last}.bias", o + self.d)]
        return out


def _prep_layers(sd, prefix, n_layers, need_grad, keep_f32=False):
    layers = []
    for l in range(n_layers):
        rb = f"{prefix}transformer.resblocks.{l}."
        f32 = lambda k: sd[rb + k].detach().float().contiguous()  # noqa: E731
        wqkv, wo = f32("attn.in_proj_weight"), f32("attn.out_proj.weight")
        wfc, wproj = f32("mlp.c_fc.weight"), f32("mlp.c_proj.weight")
        lw = LayerWeights(
            wqkv=ops.cast_f16(wqkv), bqkv=f32("attn.in_proj_bias"), wo=ops.cast_f16(wo), bo=f32("attn.out_proj.bias"),
            wfc=ops.cast_f16(wfc), bfc=f32("mlp.c_fc.bias"), wproj=ops.cast_f16(wproj), bproj=f32("mlp.c_proj.bias"))
        if need_grad:
            lw.wqkv_t, lw.wo_t = ops.transpose_cast_f16(wqkv), ops.transpose_cast_f16(wo)
            lw.wfc_t, lw.wproj_t = ops.transpose_cast_f16(wfc), ops.transpose_cast_f16(wproj)
        if keep_f32:
            lw.w32 = dict(wqkv=wqkv, wo=wo, wfc=wfc, wproj=wproj)
        layers.append(lw)
    return layers


def prepare_visual(sd: dict, prefix: str = "visual.", need_grad: bool = False) -> TowerWeights:
    """Builds kernel-ready weights from a CLIP state_dict (key layout of TPT/clip/model.py)."""
    conv = sd[prefix + "conv1.weight"].detach().float()
    d, _, p, _ = conv.shape
    pos = sd[prefix + "positional_embedding"].detach().float().contiguous()
    L = pos.shape[0]
    n_layers = len([k for k in sd if k.startswith(prefix) and k.endswith(".attn.in_proj_weight")])
    proj = sd[prefix + "proj"].detach().float().contiguous()
    k_real = 3 * p * p
    w = TowerWeights(kind="visual", d=d, heads=d // 64, n_layers=n_layers, L=L, E=proj.shape[1], patch=p,
                     resolution=p * int(round(math.sqrt(L - 1))), k_pad=_round_up(k_real, 64))
    w.conv_w = ops.cast_f16(conv.reshape(d, k_real).contiguous(), k_pad=w.k_pad)
    w.cls = sd[prefix + "class_embedding"].detach().float().contiguous()
    w.pos, w.proj = pos, proj
    w.layers = _prep_layers(sd, prefix, n_layers, need_grad)
    w.ln_flat = torch.empty((4 * n_layers + 4) * d, dtype=torch.float32, device=conv.device)
    for key, off in w.ln_names(prefix):
        w.ln_flat[off:off + d].copy_(sd[key].detach().float())
    return w


def prepare_text(sd: dict, need_grad: bool = False, keep_f32: bool = True) -> TowerWeights:
    """keep_f32: also keep (references to) the fp32 Linear weights, which the fp32 class-feature path
    (text_features(..., precise=True)) multiplies with; no copy when the checkpoint is already fp32."""
    emb = sd["token_embedding.weight"].detach().float().contiguous()
    pos = sd["positional_embedding"].detach().float().contiguous()
    d = emb.shape[1]
    n_layers = len([k for k in sd if k.startswith("transformer.resblocks.") and k.endswith(".attn.in_proj_weight")])
    proj = sd["text_projection"].detach().float().contiguous()
    w = TowerWeights(kind="text", d=d, heads=d // 64, n_layers=n_layers, L=pos.shape[0], E=proj.shape[1],
                     has_ln_pre=False)
    w.tok_emb, w.pos, w.proj = emb, pos, proj
    w.layers = _prep_layers(sd, "", n_layers, need_grad, keep_f32=keep_f32)
    w.ln_flat = torch.empty((4 * n_layers + 2) * d, dtype=torch.float32, device=emb.device)
    for key, off in w.ln_names(""):
        w.ln_flat[off:off + d].copy_(sd[key].detach().float())
    return w


def copy_tower_(dst: TowerWeights, src: TowerWeights) -> TowerWeights:
    """In-place refresh of kernel-ready weights (same architecture): every tensor of `dst` keeps its storage, so
    workspaces, views and captured CUDA graphs that point at it stay valid."""
    if (dst.kind, dst.d, dst.n_layers, dst.L, dst.E) != (src.kind, src.d, src.n_layers, src.L, src.E):
        raise RlcfError("copy_tower_: architectures differ")
    for name in ("proj", "ln_flat", "conv_w", "cls", "pos", "tok_emb"):
        a, b = getattr(dst, name), getattr(src, name)
        if a is not None and b is not None:
            a.copy_(b)
    for la, lb in zip(dst.layers, src.layers):
        for name in ("wqkv", "bqkv", "wo", "bo", "wfc", "bfc", "wproj", "bproj", "wqkv_t", "wo_t", "wfc_t", "wproj_t"):
            a, b = getattr(la, name), getattr(lb, name)
            if a is not None and b is not None:
                a.copy_(b)
    return dst


def linear(a, wt, out, M, epilogue=EPI_F16, bias=None, resid=None, aux_in=None, aux_out=None):
    """out[:M] = epilogue(a[:M] @ wt^T).  wt [N,K]: one weight for all rows.  wt [G,N,K]: per-sample weights, the M
    rows being G equal runs (one grouped launch)."""
    if wt.dim() == 2:
        return ops.gemm(a, wt, out, epilogue=epilogue, bias=bias, resid=resid, aux_in=aux_in, aux_out=aux_out, M=M)
    G = wt.shape[0]
    if M % G:
        raise RlcfError(f"{M} rows do not split into {G} weight groups")
    m = M // G

    def v(t):
        return None if t is None else t[:M].view(G, m, t.shape[-1])
    return ops.gemm_grouped(v(a), wt, v(out), epilogue=epilogue, bias=bias, resid=v(resid), aux_in=v(aux_in),
                            aux_out=v(aux_out))


@dataclass
class EmbedRows:
    """Text-tower input when every sample owns its embedding rows (text->image retrieval tunes token_embedding and
    positional_embedding, retrieval/custom_models.py:144-152): sample g's L token rows start at tok + g*stride, its
    positional embedding at pos + g*stride (floats)."""
    tok: torch.Tensor
    pos: torch.Tensor
    stride: int


@dataclass
class PromptLayout:
    """Where the learnable vectors sit inside the class prompts (PromptLearner.forward, custom_clip.py:198-289) when it
    is not the default [SOS | ctx | class tokens ...]: class token in the middle / at the front of the context, learned
    class tokens.  Learnable vectors of one parameter set: n_ctx context vectors, then (cls_pos given) one per class."""
    src_map: torch.Tensor          # int32 [C, L]: >= 0 position in the class's own token row, < 0 learnable vector -1-m
    ctx_pos: torch.Tensor          # int32 [C, n_ctx]: token position of context vector v in class c
    cls_pos: torch.Tensor | None   # int32 [C]: token position of class c's learnable class vector
    n_ctx: int

    @property
    def n_vec(self):
        return self.n_ctx + (0 if self.cls_pos is None else self.cls_pos.numel())


class ActStore:
    """Activations one training-mode forward keeps for the backward (per layer, for `rows` token rows)."""

    def __init__(self, w: TowerWeights, n_seq: int, device, full: bool = False):
        """full=True also keeps the fp16 inputs of the four linear layers (needed for weight gradients)."""
        d, L, nl = w.d, w.L, w.n_layers
        rows = n_seq * L
        f32 = dict(dtype=torch.float32, device=device)
        f16 = dict(dtype=torch.float16, device=device)
        self.n_seq, self.rows = n_seq, rows
        self.x_pre = torch.empty(rows, d, **f32) if w.has_ln_pre else None
        self.x_in = torch.empty(nl + 1, rows, d, **f32)   # x_in[l] = input of block l; x_in[nl] = tower output
        self.x_mid = torch.empty(nl, rows, d, **f32)      # after the attention residual
        self.qkv = torch.empty(nl, rows, 3 * d, **f16)
        self.attn = torch.empty(nl, rows, d, **f16)
        self.lse = torch.empty(nl, n_seq, w.heads, L, **f32)
        self.u = torch.empty(nl, rows, 4 * d, **f16)      # QuickGELU pre-activation
        self.a1 = torch.empty(nl, rows, d, **f16) if full else None      # ln_1 output (input of in_proj)
        self.a2 = torch.empty(nl, rows, d, **f16) if full else None      # ln_2 output (input of c_fc)
        self.h = torch.empty(nl, rows, 4 * d, **f16) if full else None   # QuickGELU output (input of c_proj)


class ViewStore:
    """Per-layer activations of the all-views inference pass for a chunk of images, blocks 0 .. n_layers - 2 (the last
    block of that pass only computes the class-token rows).  After the confident views are known, their slices are
    lifted into the ActStore of the backward (TowerRunner.adopt) instead of running those views a second time: the
    reference gets the same from autograd, which keeps the graph of all 64 views (tpt_cls_rl.py:55-71).  The QuickGELU
    pre-activation -- a third of the bytes -- is not kept; TowerRunner.complete recomputes it for the selected views."""

    def __init__(self, w: TowerWeights, n_seq: int, device):
        d, L, nl = w.d, w.L, w.n_layers - 1
        rows = n_seq * L
        f32 = dict(dtype=torch.float32, device=device)
        f16 = dict(dtype=torch.float16, device=device)
        self.n_seq, self.rows, self.n_layers = n_seq, rows, nl
        self.x_pre = torch.empty(rows, d, **f32) if w.has_ln_pre else None
        self.x_in = torch.empty(nl + 1, rows, d, **f32)   # x_in[nl] = input of the last block
        self.x_mid = torch.empty(nl, rows, d, **f32)
        self.qkv = torch.empty(nl, rows, 3 * d, **f16)
        self.attn = torch.empty(nl, rows, d, **f16)
        self.lse = torch.empty(nl, n_seq, w.heads, L, **f32)
        self.u = self.a1 = self.a2 = self.h = None

    @staticmethod
    def bytes_per_seq(w: TowerWeights) -> int:
        nl = w.n_layers - 1
        return w.L * w.d * (4 * (nl + 1 + nl + (1 if w.has_ln_pre else 0)) + 2 * 4 * nl) + 4 * nl * w.heads * w.L


class TowerRunner:
    """Runs one tower's forward (and LayerNorm-parameter backward) on preallocated workspaces."""

    def __init__(self, w: TowerWeights, max_seq: int):
        self.w = w
        self.max_seq = max_seq
        dev = w.ln_flat.device
        d, L = w.d, w.L
        rows = max_seq * L
        f32 = dict(dtype=torch.float32, device=dev)
        f16 = dict(dtype=torch.float16, device=dev)
        self.x = torch.empty(rows, d, **f32)
        self.a = torch.empty(rows, d, **f16)        # LayerNorm output / attention output
        self.qkv = torch.empty(rows, 3 * d, **f16)
        self.h = torch.empty(rows, 4 * d, **f16)
        if w.kind == "visual":
            self.patches = torch.empty(max_seq * (L - 1), w.k_pad, **f16)
            self.patch_out = torch.empty(max_seq * (L - 1), d, **f32)
        # Inference forwards of a vision tower run the LAST block's attention / out_proj / ln_2 / MLP on the class-token
        # rows only: nothing else of that block reaches ln_post (model.py:232-238).  Same values for those rows, 1/L of
        # the block's work after the QKV GEMM (~6 % of a 12-layer forward).  RLCF_PRUNE_LAST=0 runs every row.
        self.infer_row_stride = L
        if w.kind == "visual" and PRUNE_LAST and L <= 672 and w.n_layers > 0:
            self.infer_row_stride = 1
            self.c_x = torch.empty(3, max_seq, d, **f32)       # class-token rows: block input, after attention, output
            self.c_a = torch.empty(max_seq, d, **f16)
            self.c_q = torch.empty(max_seq, d, **f16)          # queries of the class-token rows
            self.c_h = torch.empty(max_seq, 4 * d, **f16)
            self.cls_rows = (torch.arange(max_seq, device=dev, dtype=torch.int32) * L).contiguous()
        self._out, self._out_stride = None, L
        # backward workspaces are allocated lazily by reserve_backward()
        self.g16 = self.gh = self.gqkv = self.dres = self.dres16 = None

    def reserve_backward(self, n_seq: int):
        w = self.w
        dev = w.ln_flat.device
        rows = n_seq * w.L
        if self.dres is not None and self.dres.shape[0] >= rows:
            return
        f16 = dict(dtype=torch.float16, device=dev)
        self.dres = torch.empty(rows, w.d, dtype=torch.float32, device=dev)
        self.dres16 = torch.empty(rows, w.d, **f16)
        self.g16 = torch.empty(rows, w.d, **f16)
        self.gh = torch.empty(rows, 4 * w.d, **f16)
        self.gqkv = torch.empty(rows, 3 * w.d, **f16)

    # ------------------------------------------------------------------ forward
    def _embed_visual(self, images, view_idx, n_seq, ln, pstride, rows_per_set, store, w=None):
        w = self.w if w is None else w
        P = w.L - 1
        if images.shape[-1] != w.resolution or images.shape[-2] != w.resolution:
            raise RlcfError(f"expected {w.resolution}x{w.resolution} input, got {tuple(images.shape)}")
        ops.im2col(images, view_idx, n_seq, w.patch, w.k_pad, self.patches)
        linear(self.patches, w.conv_w, self.patch_out, n_seq * P, epilogue=EPI_F32)
        x = store.x_in[0] if store is not None else self.x
        ops.embed_lnpre(self.patch_out, w.cls, w.pos, ln, ln[w.d:], pstride, rows_per_set, n_seq, w.L, w.d, x,
                        x_pre=None if store is None else store.x_pre, embed_stride=w.embed_stride)
        return x

    def forward(self, n_seq, ln, pstride=0, seqs_per_set=None, images=None, view_idx=None, tokens=None, store=None,
                causal=None, prompt=None, w=None):
        """Returns the fp32 residual stream after the last block ([n_seq*L, d]).

        ln: flat LayerNorm parameters, [P] (pstride 0) or [n_sets, P] (pstride P, seqs_per_set sequences per set).
        store: ActStore to keep activations for backward() (training-mode forward), else None.
        w: weights to use instead of the runner's own (same architecture; per-image weights in full tuning).
        """
        w = self.w if w is None else w
        if n_seq > self.max_seq:
            raise RlcfError(f"n_seq {n_seq} exceeds reserved {self.max_seq}")
        d, L = w.d, w.L
        rows = n_seq * L
        rows_per_set = rows if seqs_per_set is None else seqs_per_set * L
        lnv = ln.view(-1)
        causal = (w.kind == "text") if causal is None else causal

        def gb(off):  # gamma, beta views at a flat offset
            return lnv[off:], lnv[off + d:]

        if w.kind == "visual":
            x = self._embed_visual(images, view_idx, n_seq, lnv, pstride, rows_per_set, store, w)
        else:
            x = store.x_in[0] if store is not None else self.x
            if isinstance(prompt, EmbedRows):   # per-sample token rows + per-sample positional embedding
                ops.add_rows(prompt.tok, prompt.stride, prompt.pos, prompt.stride, n_seq, w.L * w.d, x)
            elif isinstance(prompt, torch.Tensor):   # ready-made prompt embeddings [n_seq, L, d] (TextEncoder.forward)
                x[:n_seq * w.L].copy_((prompt.float() + w.pos).reshape(n_seq * w.L, w.d))
            elif prompt is not None:   # learnable context vectors spliced into the class prompts (PromptLearner)
                ctx, ctx_stride, n_ctx, n_sets = prompt[:4]
                layout = prompt[4] if len(prompt) > 4 else None
                if layout is None:
                    ops.embed_prompts(tokens, w.tok_emb, w.pos, ctx, ctx_stride, n_ctx, n_sets, x)
                else:
                    ops.embed_prompts_map(tokens, w.tok_emb, w.pos, ctx, ctx_stride, layout.src_map, n_sets, x)
            else:
                ops.embed_text(tokens, w.tok_emb, w.pos, x)
        views = isinstance(store, ViewStore)
        if views and not (w is self.w and self.infer_row_stride == 1 and not causal):
            raise RlcfError("a ViewStore needs the class-token-only last block (vision tower, shared weights)")
        prune = (store is None or views) and w is self.w and self.infer_row_stride == 1 and not causal
        self._out_stride = L
        for l, lw in enumerate(w.layers):
            if prune and l == w.n_layers - 1:
                # last block, class-token rows only (K and V of every token are in qkv)
                g, b = gb(w.ln_off("ln_1", l))
                ops.layernorm_fwd(x, g, b, rows, d, out16=self.a, param_stride=pstride, rows_per_set=rows_per_set)
                # in_proj: keys and values of every token, the query of the class token only
                kv = self.qkv.view(-1)[:rows * 2 * d].view(rows, 2 * d)
                linear(self.a, lw.wqkv[d:], kv, rows, epilogue=EPI_F16, bias=lw.bqkv[d:])
                ops.gather_seqs(self.a, self.cls_rows, self.c_a, n_seq, 1)
                linear(self.c_a, lw.wqkv[:d], self.c_q, n_seq, epilogue=EPI_F16, bias=lw.bqkv[:d])
                cx, cmid, cout = self.c_x[0], self.c_x[1], self.c_x[2]
                sets = n_seq if seqs_per_set is None else seqs_per_set
                ops.attention_row_fwd(kv, n_seq, L, w.heads, self.c_a, q_row=0, x=x, x_row=cx, q_rows=self.c_q)
                linear(self.c_a, lw.wo, cmid, n_seq, epilogue=EPI_RESID_F32, bias=lw.bo, resid=cx)
                g, b = gb(w.ln_off("ln_2", l))
                ops.layernorm_fwd(cmid, g, b, n_seq, d, out16=self.c_a, param_stride=pstride, rows_per_set=sets)
                linear(self.c_a, lw.wfc, self.c_h, n_seq, epilogue=EPI_GELU_F16, bias=lw.bfc)
                linear(self.c_h, lw.wproj, cout, n_seq, epilogue=EPI_RESID_F32, bias=lw.bproj, resid=cmid)
                x = cout
                self._out_stride = 1
                break
            x = self._block(l, x, n_seq, lnv, pstride, rows_per_set, store, causal, w)
        self._out = x
        return x

    def _block(self, l, x, n_seq, lnv, pstride, rows_per_set, store, causal, w):
        """One residual attention block (model.py:171-192) on every row of x; returns the block's output."""
        d, L = w.d, w.L
        rows = n_seq * L
        lw = w.layers[l]
        qkv = store.qkv[l] if store is not None else self.qkv
        attn = store.attn[l] if store is not None else self.a
        full = store is not None and store.a1 is not None
        a1 = store.a1[l] if full else self.a
        off = w.ln_off("ln_1", l)
        ops.layernorm_fwd(x, lnv[off:], lnv[off + d:], rows, d, out16=a1, param_stride=pstride, rows_per_set=rows_per_set)
        linear(a1, lw.wqkv, qkv, rows, epilogue=EPI_F16, bias=lw.bqkv)
        ops.attention_fwd(qkv, n_seq, L, w.heads, attn, causal=causal, lse=None if store is None else store.lse[l])
        x_mid = store.x_mid[l] if store is not None else x
        linear(attn, lw.wo, x_mid, rows, epilogue=EPI_RESID_F32, bias=lw.bo, resid=x)
        a2 = store.a2[l] if full else self.a
        h = store.h[l] if full else self.h
        off = w.ln_off("ln_2", l)
        ops.layernorm_fwd(x_mid, lnv[off:], lnv[off + d:], rows, d, out16=a2, param_stride=pstride,
                          rows_per_set=rows_per_set)
        linear(a2, lw.wfc, h, rows, epilogue=EPI_GELU_F16, bias=lw.bfc,
               aux_out=None if store is None or store.u is None else store.u[l])
        x_next = store.x_in[l + 1] if store is not None else x
        linear(h, lw.wproj, x_next, rows, epilogue=EPI_RESID_F32, bias=lw.bproj, resid=x_mid)
        return x_next

    def adopt(self, views: ViewStore, idx: torch.Tensor, n: int, store: ActStore, seq0: int = 0):
        """Lifts the activations of sequences idx[0:n] of a ViewStore (blocks 0 .. n_layers - 2 and the input of the last
        block) into sequences seq0 .. seq0 + n of the backward's ActStore."""
        L, nl = self.w.L, views.n_layers
        if self.w.has_ln_pre:
            ops.gather_seqs(views.x_pre, idx, store.x_pre, n, L, seq0)
        ops.gather_seqs(views.x_in, idx, store.x_in[:nl + 1], n, L, seq0)
        ops.gather_seqs(views.x_mid, idx, store.x_mid[:nl], n, L, seq0)
        ops.gather_seqs(views.qkv, idx, store.qkv[:nl], n, L, seq0)
        ops.gather_seqs(views.attn, idx, store.attn[:nl], n, L, seq0)
        ops.gather_seqs(views.lse.view(nl, views.n_seq, -1), idx, store.lse[:nl].view(nl, store.n_seq, -1), n, 1, seq0)

    def complete(self, store: ActStore, n_seq, ln, pstride=0, seqs_per_set=None, images=None, view_idx=None):
        """What an adopted ActStore still lacks for the backward: the QuickGELU pre-activations of blocks 0 .. n - 2
        (ln_2 + c_fc on the stored post-attention residual) and the whole last block, run in training mode from its
        stored input.  A store that also keeps the Linear inputs for weight gradients (full tuning) gets those too:
        the ln_1 / ln_2 outputs, the QuickGELU outputs and (images, view_idx given) the conv1 patches of its
        sequences.  Returns the tower output rows like a training-mode forward()."""
        w = self.w
        d, L, n = w.d, w.L, w.n_layers
        rows = n_seq * L
        rows_per_set = rows if seqs_per_set is None else seqs_per_set * L
        lnv = ln.view(-1)
        full = store.a1 is not None
        if full and images is not None:
            ops.im2col(images, view_idx, n_seq, w.patch, w.k_pad, self.patches)
        for l in range(n - 1):
            if full:
                off = w.ln_off("ln_1", l)
                ops.layernorm_fwd(store.x_in[l], lnv[off:], lnv[off + d:], rows, d, out16=store.a1[l],
                                  param_stride=pstride, rows_per_set=rows_per_set)
            a2 = store.a2[l] if full else self.a
            off = w.ln_off("ln_2", l)
            ops.layernorm_fwd(store.x_mid[l], lnv[off:], lnv[off + d:], rows, d, out16=a2, param_stride=pstride,
                              rows_per_set=rows_per_set)
            linear(a2, w.layers[l].wfc, store.h[l] if full else self.h, rows, epilogue=EPI_GELU_F16,
                   bias=w.layers[l].bfc, aux_out=store.u[l])
        x = self._block(n - 1, store.x_in[n - 1], n_seq, lnv, pstride, rows_per_set, store, False, w)
        self._out, self._out_stride = x, L
        return x

    def head(self, x, n_seq, ln, pstride=0, seqs_per_set=None, row_idx=None, class_feat=None, logit_scale=1.0,
             feat=None, inv_norm=None, logits=None, w=None):
        """ln_post/ln_final on one row per sequence -> projection -> L2 normalise (-> logits)."""
        w = self.w if w is None else w
        lnv = ln.view(-1)
        off = w.ln_off("ln_post")
        # x is either [n_seq*L, d] (one row of each sequence is used) or the class-token rows [n_seq, d] of a pruned forward
        row_stride = self._out_stride if x is self._out else w.L
        ops.head_fwd(x, lnv[off:], lnv[off + w.d:], w.proj, n_seq, w.d, w.E, feat=feat, inv_norm=inv_norm,
                     logits=logits, class_feat=class_feat, logit_scale=logit_scale, row_idx=row_idx, row_stride=row_stride,
                     param_stride=pstride, seqs_per_set=seqs_per_set, proj_stride=w.proj_stride)

    # ------------------------------------------------------------------ backward (LayerNorm parameters only)
    def backward(self, store: ActStore, n_sets, seqs_per_set, ln, pstride, partials, n_slots=N_SLOTS, w=None,
                 hook=None):
        """Propagates self.dres (gradient w.r.t. the tower output rows, fp32, already filled by head_bwd) down to
        ln_pre, writing every LayerNorm's d(gamma), d(beta) partials.  Without `hook` the GEMM weights are frozen
        (dgrad only); with it, hook.linear(layer, name, dY, X) is called where a weight gradient dY^T X is due and
        hook.embed(dx_pre) at the end (full image-encoder tuning)."""
        w = self.w if w is None else w
        d, L = w.d, w.L
        n_seq = n_sets * seqs_per_set
        rows = n_seq * L
        rows_per_set = seqs_per_set * L
        lnv = ln.view(-1)
        P = w.P
        dres, dres16 = self.dres, self.dres16
        ops.cast_f16(dres, out=dres16, rows=rows)
        for l in range(w.n_layers - 1, -1, -1):
            lw = w.layers[l]
            # MLP branch: d u = (d x_out @ Wproj) * gelu'(u);  d a2 = d u @ Wfc
            if hook is not None:
                hook.linear(l, "c_proj", dres16, store.h[l])
            linear(dres16, lw.wproj_t, self.gh, rows, epilogue=EPI_GELU_BWD_F16, aux_in=store.u[l])
            if hook is not None:
                hook.linear(l, "c_fc", self.gh, store.a2[l])
            linear(self.gh, lw.wfc_t, self.g16, rows, epilogue=EPI_F16)
            off = w.ln_off("ln_2", l)
            ops.layernorm_bwd(self.g16, store.x_mid[l], lnv[off:], rows_per_set, n_sets, d, partials, n_slots, P, off,
                              dx=dres, accumulate=True, param_stride=pstride, dx16=dres16)
            # attention branch
            if hook is not None:
                hook.linear(l, "out_proj", dres16, store.attn[l])
            linear(dres16, lw.wo_t, self.g16, rows, epilogue=EPI_F16)
            ops.attention_bwd(store.qkv[l], store.attn[l], self.g16, store.lse[l], n_seq, L, w.heads, self.gqkv,
                              causal=(w.kind == "text"))
            if hook is not None:
                hook.linear(l, "in_proj", self.gqkv, store.a1[l])
            linear(self.gqkv, lw.wqkv_t, self.g16, rows, epilogue=EPI_F16)
            off = w.ln_off("ln_1", l)
            ops.layernorm_bwd(self.g16, store.x_in[l], lnv[off:], rows_per_set, n_sets, d, partials, n_slots, P, off,
                              dx=dres, accumulate=True, param_stride=pstride, dx16=dres16)
        if w.has_ln_pre:
            dxp = hook.dx_pre if hook is not None else None
            ops.layernorm_bwd(dres, store.x_pre, lnv, rows_per_set, n_sets, d, partials, n_slots, P, 0, dx=dxp,
                              accumulate=False, param_stride=pstride)
            if hook is not None:
                hook.embed(dxp)


@dataclass
class RlcfConfig:
    """Hyper-parameters of the TTA loop; names follow TPT/params.py:23-73."""
    n_views: int = 64            # --batch_size (views per test image, first one is the clean view)
    selection_p: float = 0.1     # --selection_p
    tta_steps: int = 1           # --tta_steps
    sample_k: int = 3            # --sample_k
    lr: float = 5e-3             # --lr
    weight_decay: float = 5e-4   # --weight_decay
    betas: tuple = (0.9, 0.999)
    eps: float = 1e-8
    clipscore_weight: float = 2.5
    reward_process: bool = True  # --reward_process
    process_batch: bool = False  # --process_batch
    reward_amplify: bool = False  # --reward_amplify
    loss_scale: float = 1024.0   # static gradient scale for the fp16 dgrad operands (the reference's GradScaler(1000))
    loss: str = "rlcf"           # "rlcf" (tpt_cls_rl.py:63-71) | "tpt" (avg_entropy, tpt_cls_rl.py:38-44)
    reward_weights: tuple = ()   # ensemble of reward models (CLIPRewardsMultiple, clip_reward.py:180-307): one weight
                                 # per model -- normalised confidences, or 1/n each for weighted_scores = False
    min_entropy_w: float = 0.0   # --min_entropy_reg 1 --min_entropy_w w: loss += w * avg_entropy(output)
                                 # (tpt_cls_rl.py:73-74); 0 = off

    @property
    def n_selected(self):
        return int(self.n_views * self.selection_p)  # tpt_cls_rl.py:34


def setup_view_store(eng, policy: TowerWeights, run: "TowerRunner", B: int, V: int, S: int, dev):
    """The all-views pass keeps its per-layer activations for a chunk of images (ViewStore) so that the selected views
    need no second, training-mode forward: their slices are lifted into eng.store (autograd does the same for the
    reference by keeping the graph of all 64 views).  1.7 GB per image at ViT-B/16 x 64 views, so the images of a step go
    through in chunks that fit RLCF_VIEW_STORE_GB (default 48; 0 = run the selected views twice, as in round 1) and
    what the device has free."""
    eng.views, eng.view_chunk, eng.view_store_gb = None, 0, 0.0
    budget = float(os.environ.get("RLCF_VIEW_STORE_GB", "48")) * 2 ** 30
    per_img = V * ViewStore.bytes_per_seq(policy)
    if dev.type == "cuda":
        budget = min(budget, torch.cuda.mem_get_info(dev)[0] - 16 * 2 ** 30)     # leave room for the other workspaces
    if run.infer_row_stride == 1 and policy.n_layers > 1 and budget >= per_img:
        n_chunks = -(-B // max(1, min(B, int(budget // per_img))))
        eng.view_chunk = -(-B // n_chunks)
        try:
            eng.views = ViewStore(policy, eng.view_chunk * V, dev)
        except torch.OutOfMemoryError:      # fragmented or shared device: the second-forward route needs no store
            eng.views, eng.view_chunk = None, 0
            return
        eng.view_store_gb = eng.view_chunk * per_img / 2 ** 30
        eng.sel_local = torch.empty(B * S, dtype=torch.int32, device=dev)
        eng.chunk_off = (torch.arange(B, device=dev, dtype=torch.int32) // eng.view_chunk * (eng.view_chunk * V)
                         ).repeat_interleave(S).contiguous()


def all_views_pass(eng, run: "TowerRunner", init_ln: torch.Tensor, images: torch.Tensor, B: int, V: int, S: int, C: int):
    """Forward of all B x V views with the shared initial parameters, logits, entropy selection -- and, with a ViewStore,
    adoption of the selected views' activations into eng.store, chunk of images by chunk."""
    if eng.views is None:
        x = run.forward(B * V, init_ln, images=images)
        run.head(x, B * V, init_ln, class_feat=eng.class_feat, logit_scale=eng.logit_scale, logits=eng.logits_all)
        ops.entropy_select(eng.logits_all, B, V, C, S, eng.sel, eng.sel_global, eng.entropy)
        return
    for c0 in range(0, B, eng.view_chunk):
        n = min(eng.view_chunk, B - c0)
        x = run.forward(n * V, init_ln, images=images[c0 * V:(c0 + n) * V], store=eng.views)
        run.head(x, n * V, init_ln, class_feat=eng.class_feat, logit_scale=eng.logit_scale,
                 logits=eng.logits_all[c0 * V:(c0 + n) * V])
        loc = eng.sel_local[c0 * S:(c0 + n) * S]                  # view numbers inside the chunk
        ops.entropy_select(eng.logits_all[c0 * V:(c0 + n) * V], n, V, C, S, eng.sel[c0:c0 + n], loc,
                           eng.entropy[c0:c0 + n])
        run.adopt(eng.views, loc, n * S, eng.store, seq0=c0 * S)
    torch.add(eng.sel_local, eng.chunk_off, out=eng.sel_global)    # view numbers inside the step's batch


class RewardScorer:
    """The frozen reward model(s) of one engine: image features of the selected views and the reward-weighted loss.
    One tower (CLIPRewards, clip_reward.py:43-178) or up to four (CLIPRewardsMultiple, clip_reward.py:180-307), each
    with its class features [C, E_i]; `weights` as RlcfConfig.reward_weights."""

    def __init__(self, reward, class_feat, n_seq_max: int, weights=()):
        self.towers = list(reward) if isinstance(reward, (list, tuple)) else [reward]
        cls = list(class_feat) if isinstance(class_feat, (list, tuple)) else [class_feat]
        if len(cls) != len(self.towers) or not 1 <= len(self.towers) <= 4:
            raise RlcfError("one set of reward class features per reward tower (1..4 towers)")
        self.class_feats = [c.float().contiguous() for c in cls]
        for t, c in zip(self.towers, self.class_feats):
            if c.shape[1] != t.E:
                raise RlcfError(f"reward class features have width {c.shape[1]}, the tower embeds to {t.E}")
        self.weights = [float(x) for x in weights] if len(self.towers) > 1 else [1.0]
        if len(self.weights) != len(self.towers):
            raise RlcfError("RlcfConfig.reward_weights must hold one weight per reward model")
        dev = self.towers[0].ln_flat.device
        self.runners = [TowerRunner(t, n_seq_max) for t in self.towers]
        self.feats = [torch.empty(n_seq_max, t.E, dtype=torch.float32, device=dev) for t in self.towers]
        self.resized = [None] * len(self.towers)    # views at the tower's own resolution, allocated on first use

    def features(self, images, view_idx, n_seq):
        """reward_model.set_image_features(inputs[selected_idx])          (tpt_cls_rl.py:59, clip_reward.py:130-137),
        with the bicubic resize of clip_reward.py:133-134 when a reward model runs at another resolution."""
        for i, (t, r, f) in enumerate(zip(self.towers, self.runners, self.feats)):
            if images.shape[-1] != t.resolution or images.shape[-2] != t.resolution:
                if self.resized[i] is None:
                    self.resized[i] = torch.empty(r.max_seq, images.shape[1], t.resolution, t.resolution,
                                                  dtype=torch.float32, device=images.device)
                ops.bicubic_resize(images, view_idx, n_seq, self.resized[i])
                x = r.forward(n_seq, t.ln_flat, images=self.resized[i])
            else:
                x = r.forward(n_seq, t.ln_flat, images=images, view_idx=view_idx)
            r.head(x, n_seq, t.ln_flat, feat=f)

    def loss(self, logits, n_img, S, K, C, dlogits, cfg, **outs):
        kw = dict(clipscore_weight=cfg.clipscore_weight, reward_process=cfg.reward_process,
                  process_batch=cfg.process_batch, amplify=cfg.reward_amplify, loss_scale=cfg.loss_scale, **outs)
        if len(self.towers) == 1:
            ops.reward_loss(logits, None, self.feats[0], self.class_feats[0], n_img, S, K, C, dlogits, **kw)
        else:
            ops.reward_loss_multi(logits, None, self.feats, self.class_feats, self.weights, n_img, S, K, C, dlogits, **kw)

    def fwd_flops(self) -> float:
        return float(sum(RlcfEngine.tower_fwd_flops(t, cls_only_last=True) for t in self.towers))


class RlcfEngine:
    """Batched LN-only RLCF adaptation (TPT/tune_cls_rl.py --tune_norm 1): `n_img` independent test images per call.

    Every image restarts from the same initial LayerNorm parameters and an empty Adam state
    (tune_cls_rl.py:210-213), so images are independent and can share the frozen GEMM weights in one launch
    sequence; each owns a private [P] slice of LayerNorm parameters, Adam moments and gradient partials.
    """

    def __init__(self, policy: TowerWeights, class_feat: torch.Tensor, logit_scale: float, cfg: RlcfConfig,
                 n_img: int, reward: TowerWeights | None = None, reward_class_feat: torch.Tensor | None = None):
        if policy.layers[0].wqkv_t is None:
            raise RlcfError("policy weights must be prepared with need_grad=True")
        self.cfg, self.n_img = cfg, n_img
        self.policy, self.reward = policy, reward
        self.class_feat = class_feat.float().contiguous()
        self.logit_scale = float(logit_scale)
        dev = policy.ln_flat.device
        V, S, C = cfg.n_views, cfg.n_selected, self.class_feat.shape[0]
        if S < 1:
            raise RlcfError(f"int(n_views * selection_p) = {S}: no view would be selected (tpt_cls_rl.py:34)")
        if cfg.loss == "rlcf" and (reward is None or reward_class_feat is None):
            raise RlcfError("RLCF loss needs a reward tower and reward class features")
        B, P = n_img, policy.P
        self.run = TowerRunner(policy, B * V)
        self.run.reserve_backward(B * S)
        self.store = ActStore(policy, B * S, dev)
        self.scorer = RewardScorer(reward, reward_class_feat, B * S, cfg.reward_weights) if reward is not None else None
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        setup_view_store(self, policy, self.run, B, V, S, dev)
        self.init_params = policy.ln_flat.clone()
        self.params = torch.empty(B, P, **f32)
        self.m = torch.empty(B, P, **f32)
        self.v = torch.empty(B, P, **f32)
        self.n_slots = max(N_SLOTS, S)   # head_bwd writes one slot per selected view
        self.partials = torch.empty(B, self.n_slots, P, **f32)
        self.grad = torch.empty(B, P, **f32)
        self.logits_all = torch.empty(B * V, C, **f32)
        self.entropy = torch.empty(B, V, **f32)
        self.sel = torch.empty(B, S, **i32)
        self.sel_global = torch.empty(B * S, **i32)
        self.first_view = (torch.arange(B, device=dev, dtype=torch.int32) * V).contiguous()
        self.logits_sel = torch.empty(B * S, C, **f32)
        self.feat_sel = torch.empty(B * S, policy.E, **f32)
        self.inv_norm_sel = torch.empty(B * S, **f32)
        self.dlogits = torch.empty(B * S, C, **f32)
        self.topk_idx = torch.empty(B * S, cfg.sample_k, **i32)
        self.scores = torch.empty(B * S, cfg.sample_k, **f32)
        self.rewards = torch.empty(B * S, cfg.sample_k, **f32)
        self.loss = torch.empty(cfg.tta_steps, B, **f32)
        self.logits_final = torch.empty(B, C, **f32)
        self._graph = None
        self._static_images = None

    @property
    def reward_feat(self):
        """Reward-model image features of the selected views (a list for an ensemble of reward models)."""
        if self.scorer is None:
            return None
        return self.scorer.feats[0] if len(self.scorer.feats) == 1 else self.scorer.feats

    # ------------------------------------------------------------------
    def adapt(self, images: torch.Tensor) -> torch.Tensor:
        """images: fp32 [n_img * n_views, 3, H, W] (view 0 of each image is the clean view).
        Runs reset -> tta_steps x (select, sample, reward, weighted-CE backward, AdamW) -> adapted 1-view logits.
        Returns logits_final [n_img, C] (a workspace tensor that the next call overwrites)."""
        self.tune(images)
        return self.predict(images)

    def predict(self, images: torch.Tensor) -> torch.Tensor:
        """Adapted prediction on the clean view of every image (tune_cls_rl.py:218-222) with its own parameters."""
        B, P = self.n_img, self.policy.P
        xf = self.run.forward(B, self.params, pstride=P, seqs_per_set=1, images=images, view_idx=self.first_view)
        self.run.head(xf, B, self.params, pstride=P, seqs_per_set=1, class_feat=self.class_feat,
                      logit_scale=self.logit_scale, logits=self.logits_final)
        return self.logits_final

    def tune(self, images: torch.Tensor) -> torch.Tensor:
        """test_time_tuning (tpt_cls_rl.py:47-79) for n_img images at once, starting from init_params and an empty
        Adam state.  Leaves the adapted LayerNorm slices in self.params [n_img, P] and returns them."""
        cfg, B = self.cfg, self.n_img
        V, S, K, C = cfg.n_views, cfg.n_selected, cfg.sample_k, self.class_feat.shape[0]
        pol, P = self.policy, self.policy.P
        if images.shape[0] != B * V:
            raise RlcfError(f"expected {B * V} views, got {images.shape[0]}")
        # model.reset(); optimizer.load_state_dict(optim_state)      (tune_cls_rl.py:210-213)
        ops.reset_params(self.init_params, self.params, self.m, self.v, B, P)
        # step 0: all views with the shared initial parameters (tpt_cls_rl.py:57), select_confident_samples (:58)
        all_views_pass(self, self.run, self.init_params, images, B, V, S, C)
        # reward_model.set_image_features(inputs[selected_idx])       (tpt_cls_rl.py:59)
        if cfg.loss == "rlcf":
            self.scorer.features(images, self.sel_global, B * S)
        for step in range(1, cfg.tta_steps + 1):
            if step == 1 and self.views is not None:
                # the parameters are still the initial ones: the adopted activations ARE this forward's; only the
                # QuickGELU pre-activations and the last block (class-token rows only in the all-views pass) are run
                xs = self.run.complete(self.store, B * S, self.params, pstride=P, seqs_per_set=S)
            else:
                # training-mode forward of the selected views with each image's own parameters (tpt_cls_rl.py:55)
                xs = self.run.forward(B * S, self.params, pstride=P, seqs_per_set=S, images=images,
                                      view_idx=self.sel_global, store=self.store)
            self.run.head(xs, B * S, self.params, pstride=P, seqs_per_set=S, class_feat=self.class_feat,
                          logit_scale=self.logit_scale, feat=self.feat_sel, inv_norm=self.inv_norm_sel,
                          logits=self.logits_sel)
            if cfg.loss == "rlcf":
                self.scorer.loss(self.logits_sel, B, S, K, C, self.dlogits, cfg, topk_idx=self.topk_idx,
                                 scores=self.scores, rewards=self.rewards, loss=self.loss[step - 1])
                if cfg.min_entropy_w:                                                  # tpt_cls_rl.py:73-74
                    ops.avg_entropy_reg(self.logits_sel, None, B, S, C, self.dlogits, cfg.min_entropy_w,
                                        loss=self.loss[step - 1], loss_scale=cfg.loss_scale)
            else:
                ops.avg_entropy_loss(self.logits_sel, None, B, S, C, self.dlogits, loss=self.loss[step - 1],
                                     loss_scale=cfg.loss_scale)
            self.partials.zero_()
            self.run.dres[:B * S * pol.L].zero_()
            off = pol.ln_off("ln_post")
            ops.head_bwd(self.dlogits, xs, self.params.view(-1)[off:], pol.proj, self.class_feat, self.logit_scale,
                         self.feat_sel, self.inv_norm_sel, B, S, pol.d, pol.E, C, self.run.dres, self.partials,
                         self.n_slots, P, off, row_stride=pol.L, param_stride=P)
            self.run.backward(self.store, B, S, self.params, P, self.partials, self.n_slots)
            ops.adamw_step(self.params, self.m, self.v, self.partials, B, self.n_slots, P, cfg.lr, step,
                           beta1=cfg.betas[0], beta2=cfg.betas[1], eps=cfg.eps, weight_decay=cfg.weight_decay,
                           loss_scale=cfg.loss_scale, grad_out=self.grad)
        return self.params

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def capture(self, images_like: torch.Tensor):
        """Captures adapt() into a CUDA graph reading from a static input buffer (launch-bound otherwise:
        ~450 kernels per step)."""
        self._static_images = torch.empty_like(images_like)
        self._static_images.copy_(images_like)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):  # warm-up outside capture (lazy cudaFuncSetAttribute, tensor-map entry point)
                self.adapt(self._static_images)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.adapt(self._static_images)
        self._graph = g
        return g

    def adapt_graph(self, images: torch.Tensor) -> torch.Tensor:
        if self._graph is None:
            self.capture(images)
        self._static_images.copy_(images, non_blocking=True)
        self._graph.replay()
        return self.logits_final

    def adapt_host(self, images_pinned: torch.Tensor, out_pinned: torch.Tensor) -> torch.Tensor:
        """End-to-end call with HOST buffers: pinned fp32 views in, adapted logits [n_img, C] out (pinned).
        The H2D copy, the graph replay and the D2H copy are enqueued on the current stream; the caller
        synchronises (or keeps several calls in flight)."""
        if self._graph is None:
            self.capture(images_pinned.to(self.logits_final.device))
        self._static_images.copy_(images_pinned, non_blocking=True)
        self._graph.replay()
        out_pinned.copy_(self.logits_final, non_blocking=True)
        return out_pinned

    def host_pipeline(self) -> "HostPipeline":
        """Double-buffered host->device feeding for back-to-back adapt calls (see HostPipeline)."""
        if self._graph is None:
            raise RlcfError("capture() the step first")
        return HostPipeline(self)

    # FLOP accounting (2*MACs), SURVEY.md 8(d)
    @staticmethod
    def tower_fwd_flops(w: TowerWeights, cls_only_last: bool = False) -> float:
        """cls_only_last: the inference forward of a vision tower, whose last block runs out_proj / MLP and attention for
        the class-token row only (TowerRunner.forward) -- count what is needed, not what the reference executes."""
        L, d, n, E = w.L, w.d, w.n_layers, w.E
        conv = 2 * (L - 1) * d * 3 * w.patch * w.patch if w.kind == "visual" else 0
        block = 24 * L * d * d + 4 * L * L * d
        if cls_only_last and w.kind == "visual" and PRUNE_LAST and L <= 672 and n > 0:
            last = 4 * L * d * d + 4 * L * d + 20 * d * d          # K, V of every token; one query row; one row of the rest
            return conv + (n - 1) * block + last + 2 * d * E
        return conv + n * block + 2 * d * E

    @staticmethod
    def tower_dgrad_flops(w: TowerWeights) -> float:
        L, d, n = w.L, w.d, w.n_layers
        return n * (24 * L * d * d + 8 * L * L * d)

    def algorithmic_flops_per_image(self) -> float:
        cfg = self.cfg
        V, S = cfg.n_views, cfg.n_selected
        f = self.tower_fwd_flops(self.policy)
        fi = self.tower_fwd_flops(self.policy, cls_only_last=True)      # the V-view and the final inference forwards
        total = V * fi + S * self.tower_dgrad_flops(self.policy) + fi
        total += (cfg.tta_steps - 1) * S * (f + self.tower_dgrad_flops(self.policy))
        if self.scorer is not None and cfg.loss == "rlcf":
            total += S * self.scorer.fwd_flops()
        return float(total)

    def reference_flops_per_image(self) -> float:
        """SURVEY.md 8(d)'s figure: the same count with every block run on every token, as the reference executes it."""
        cfg = self.cfg
        V, S = cfg.n_views, cfg.n_selected
        f = self.tower_fwd_flops(self.policy)
        total = V * f + S * self.tower_dgrad_flops(self.policy) + f
        total += (cfg.tta_steps - 1) * S * (f + self.tower_dgrad_flops(self.policy))
        if self.scorer is not None and cfg.loss == "rlcf":
            total += S * sum(self.tower_fwd_flops(t) for t in self.scorer.towers)
        return float(total)


class PromptEngine:
    """Batched RLCF / TPT prompt tuning (TPT/tpt_cls_rl.py with ClipTestTimeTuning, custom_clip.py:292-344):
    `n_img` independent test images per call.  The image tower runs without gradient (custom_clip.py:326-327); the
    text tower is re-run over all C class prompts with each image's own context vectors, and the backward goes
    through the text tower down to those n_ctx x d vectors -- the only trainable parameters (tpt_cls_rl.py:103-120)."""

    def __init__(self, visual: TowerWeights, text: TowerWeights, tokens: torch.Tensor, ctx_init: torch.Tensor,
                 logit_scale: float, cfg: RlcfConfig, n_img: int, reward: TowerWeights | None = None,
                 reward_class_feat: torch.Tensor | None = None, layout: PromptLayout | None = None):
        """ctx_init: the learnable vectors [n_vec, d] -- the context vectors, followed (layout.cls_pos) by one learnable
        class vector per class.  layout=None is the default arrangement [SOS | ctx | class tokens, EOS]."""
        if text.layers[0].wqkv_t is None:
            raise RlcfError("text tower must be prepared with need_grad=True")
        self.layout = layout
        if layout is not None and ctx_init.shape[0] != layout.n_vec:
            raise RlcfError(f"{ctx_init.shape[0]} learnable vectors given, the layout has {layout.n_vec}")
        dev = text.ln_flat.device
        self.cfg, self.n_img, self.visual, self.text, self.reward = cfg, n_img, visual, text, reward
        self.logit_scale = float(logit_scale)
        self.tokens = tokens.to(device=dev, dtype=torch.int64).contiguous()
        # The text tower is causal (model.py:328-334) and only the EOT row of a prompt is read (model.py:352-354): the
        # zero padding behind the EOT -- 60-odd of CLIP's 77 positions for "a photo of a <class>." -- cannot reach it.
        # Run the tower on the first max(EOT) + 1 positions (rounded up to 8): same EOT rows, same context gradients,
        # a fifth of the work the reference spends there.  RLCF_TEXT_TRUNCATE=0 keeps all 77.
        L_full = self.tokens.shape[1]
        self.full_text_positions = L_full
        need = int(self.tokens.argmax(dim=-1).max()) + 1
        if layout is not None:       # source positions the assembled prefix reads from
            need = max(need, int(layout.src_map[:, :need].max()) + 1)
        L_eff = min(L_full, (need + 7) // 8 * 8)
        if TRUNCATE_TEXT and L_eff < L_full:
            import dataclasses
            text = dataclasses.replace(text, L=L_eff, pos=text.pos[:L_eff].contiguous())
            self.text = text
            self.tokens = self.tokens[:, :L_eff].contiguous()
            if layout is not None:
                layout = dataclasses.replace(layout, src_map=layout.src_map[:, :L_eff].contiguous())
                self.layout = layout
        self.text_tokens = self.tokens.shape[1]
        C, L = self.tokens.shape
        n_vec, d = ctx_init.shape
        self.n_ctx = n_vec if layout is None else layout.n_ctx
        B, V, S = n_img, cfg.n_views, cfg.n_selected
        if S < 1:
            raise RlcfError(f"int(n_views * selection_p) = {S}: no view would be selected (tpt_cls_rl.py:34)")
        if cfg.loss == "rlcf" and (reward is None or reward_class_feat is None):
            raise RlcfError("RLCF loss needs a reward tower and reward class features")
        self.P = n_vec * d
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.irun = TowerRunner(visual, B * V)
        self.trun = TowerRunner(text, B * C)
        self.trun.reserve_backward(B * C)
        self.tstore = ActStore(text, B * C, dev)
        self.scorer = RewardScorer(reward, reward_class_feat, B * S, cfg.reward_weights) if reward is not None else None
        self.init_ctx = ctx_init.detach().float().reshape(-1).contiguous().clone()
        self.ctx = torch.empty(B, self.P, **f32)
        self.m = torch.empty(B, self.P, **f32)
        self.v = torch.empty(B, self.P, **f32)
        self.dctx = torch.empty(B, 1, self.P, **f32)
        self.grad = torch.empty(B, self.P, **f32)
        eot = self.tokens.argmax(dim=-1).to(torch.int32)
        self.eot_rows = ((torch.arange(B * C, device=dev, dtype=torch.int32) * L)
                         + eot.repeat(B)).contiguous()
        self.img_feat_all = torch.empty(B * V, visual.E, **f32)
        self.img_feat_sel = torch.empty(B * S, visual.E, **f32)
        self.img_feat_final = torch.empty(B, visual.E, **f32)
        self.txt_feat0 = torch.empty(C, text.E, **f32)
        self.txt_feat = torch.empty(B * C, text.E, **f32)
        self.txt_inv = torch.empty(B * C, **f32)
        self.logits_all = torch.empty(B * V, C, **f32)
        self.entropy = torch.empty(B, V, **f32)
        self.sel = torch.empty(B, S, **i32)
        self.sel_global = torch.empty(B * S, **i32)
        self.sel_rows = torch.empty(B * S, **i32)
        self.first_view = (torch.arange(B, device=dev, dtype=torch.int32) * V).contiguous()
        self.logits_sel = torch.empty(B * S, C, **f32)
        self.dlogits = torch.empty(B * S, C, **f32)
        self.topk_idx = torch.empty(B * S, cfg.sample_k, **i32)
        self.scores = torch.empty(B * S, cfg.sample_k, **f32)
        self.rewards = torch.empty(B * S, cfg.sample_k, **f32)
        self.loss = torch.empty(cfg.tta_steps, B, **f32)
        self.logits_final = torch.empty(B, C, **f32)
        self._graph = None
        self._static_images = None
        self.refresh_initial_text_features()

    reward_feat = RlcfEngine.reward_feat

    def refresh_initial_text_features(self):
        """Text features of the un-adapted prompts (shared by every image for the step-0 logits of all views)."""
        C = self.tokens.shape[0]
        txt = self.text
        if txt.layers[0].w32 is not None:
            # once per dataset (and per prompt initialisation): fp32, like the class features of the LN-tuning path
            off = txt.ln_off("ln_final")
            for s in range(0, C, 128):
                e = min(C, s + 128)
                frun = TextRunnerF32(txt, e - s) if s == 0 or e - s != frun.max_seq else frun
                if self.layout is None:
                    ops.embed_prompts(self.tokens[s:e].contiguous(), txt.tok_emb, txt.pos, self.init_ctx, 0, self.n_ctx,
                                      1, frun.x)
                else:      # class c's own vector index is absolute (n_ctx + c): the map rows carry it
                    ops.embed_prompts_map(self.tokens[s:e].contiguous(), txt.tok_emb, txt.pos, self.init_ctx, 0,
                                          self.layout.src_map[s:e].contiguous(), 1, frun.x)
                x = frun.layers(e - s)
                rows = (self.eot_rows[s:e] - s * txt.L).contiguous()
                ops.head_fwd(x, txt.ln_flat[off:], txt.ln_flat[off + txt.d:], txt.proj, e - s, txt.d, txt.E,
                             feat=self.txt_feat0[s:e], row_idx=rows, row_stride=txt.L)
            return
        x = self.trun.forward(C, txt.ln_flat, tokens=self.tokens, prompt=(self.init_ctx, 0, self.n_ctx, 1, self.layout))
        self.trun.head(x, C, txt.ln_flat, row_idx=self.eot_rows[:C].contiguous(), feat=self.txt_feat0)

    def _text_features(self, store):
        B, C = self.n_img, self.tokens.shape[0]
        x = self.trun.forward(B * C, self.text.ln_flat, tokens=self.tokens, store=store,
                              prompt=(self.ctx, self.P, self.n_ctx, B, self.layout))
        self.trun.head(x, B * C, self.text.ln_flat, row_idx=self.eot_rows, feat=self.txt_feat, inv_norm=self.txt_inv)
        return x

    def tune(self, images: torch.Tensor) -> torch.Tensor:
        cfg, B = self.cfg, self.n_img
        V, S, K = cfg.n_views, cfg.n_selected, cfg.sample_k
        C, L = self.tokens.shape
        txt, E = self.text, self.text.E
        if images.shape[0] != B * V:
            raise RlcfError(f"expected {B * V} views, got {images.shape[0]}")
        ops.reset_params(self.init_ctx, self.ctx, self.m, self.v, B, self.P)     # model.reset() + empty Adam state
        x = self.irun.forward(B * V, self.visual.ln_flat, images=images)          # image tower, no gradient
        self.irun.head(x, B * V, self.visual.ln_flat, feat=self.img_feat_all)
        ops.pair_logits(self.img_feat_all, self.txt_feat0, 0, 1, B * V, C, E, self.logit_scale, self.logits_all)
        ops.entropy_select(self.logits_all, B, V, C, S, self.sel, self.sel_global, self.entropy)
        torch.mul(self.sel_global, self.irun.infer_row_stride, out=self.sel_rows)   # class-token row of each selected view
        self.irun.head(x, B * S, self.visual.ln_flat, row_idx=self.sel_rows, feat=self.img_feat_sel)
        if cfg.loss == "rlcf":
            self.scorer.features(images, self.sel_global, B * S)
        for step in range(1, cfg.tta_steps + 1):
            xs = self._text_features(self.tstore)
            ops.pair_logits(self.img_feat_sel, self.txt_feat, C * E, B, S, C, E, self.logit_scale, self.logits_sel)
            if cfg.loss == "rlcf":
                self.scorer.loss(self.logits_sel, B, S, K, C, self.dlogits, cfg, topk_idx=self.topk_idx,
                                 scores=self.scores, rewards=self.rewards, loss=self.loss[step - 1])
                if cfg.min_entropy_w:                                                  # tpt_cls_rl.py:73-74
                    ops.avg_entropy_reg(self.logits_sel, None, B, S, C, self.dlogits, cfg.min_entropy_w,
                                        loss=self.loss[step - 1], loss_scale=cfg.loss_scale)
            else:
                ops.avg_entropy_loss(self.logits_sel, None, B, S, C, self.dlogits, loss=self.loss[step - 1],
                                     loss_scale=cfg.loss_scale)
            self.trun.dres[:B * C * L].zero_()
            off = txt.ln_off("ln_final")
            # text-side head backward: sequences = class prompts, "classes" = this image's S selected views
            ops.head_bwd_ex(self.dlogits, (S * C, 1, C), xs, txt.ln_flat[off:], txt.proj, self.img_feat_sel, S * E,
                            self.logit_scale, self.txt_feat, self.txt_inv, B, C, txt.d, E, S, self.trun.dres,
                            row_idx=self.eot_rows)
            self.trun.backward(self.tstore, B, C, txt.ln_flat, 0, None)
            if self.layout is None:
                ops.ctx_grad(self.trun.dres, B, C, L, self.n_ctx, txt.d, self.dctx)
            else:
                ops.vec_grad_map(self.trun.dres, self.layout.ctx_pos, self.layout.cls_pos, B, C, L, self.n_ctx, txt.d,
                                 self.dctx)
            ops.adamw_step(self.ctx, self.m, self.v, self.dctx, B, 1, self.P, cfg.lr, step, beta1=cfg.betas[0],
                           beta2=cfg.betas[1], eps=cfg.eps, weight_decay=cfg.weight_decay,
                           loss_scale=cfg.loss_scale, grad_out=self.grad)
        return self.ctx

    def predict(self, images: torch.Tensor) -> torch.Tensor:
        B, C, E = self.n_img, self.tokens.shape[0], self.text.E
        xf = self.irun.forward(B, self.visual.ln_flat, images=images, view_idx=self.first_view)
        self.irun.head(xf, B, self.visual.ln_flat, feat=self.img_feat_final)
        self._text_features(None)
        ops.pair_logits(self.img_feat_final, self.txt_feat, C * E, B, 1, C, E, self.logit_scale, self.logits_final)
        return self.logits_final

    def adapt(self, images: torch.Tensor) -> torch.Tensor:
        self.tune(images)
        return self.predict(images)

    capture = RlcfEngine.capture
    adapt_graph = RlcfEngine.adapt_graph
    adapt_host = RlcfEngine.adapt_host
    host_pipeline = RlcfEngine.host_pipeline

    def algorithmic_flops_per_image(self) -> float:
        cfg, C = self.cfg, self.tokens.shape[0]
        fi, ft = RlcfEngine.tower_fwd_flops(self.visual, cls_only_last=True), RlcfEngine.tower_fwd_flops(self.text)
        total = cfg.n_views * fi + fi + C * ft + cfg.tta_steps * C * (ft + RlcfEngine.tower_dgrad_flops(self.text))
        if self.scorer is not None and cfg.loss == "rlcf":
            total += cfg.n_selected * self.scorer.fwd_flops()
        return float(total)

    def reference_flops_per_image(self) -> float:
        """The same count as the reference executes it: every block on every token of the image towers, the text tower
        on all of CLIP's context positions (the padding behind the EOT included)."""
        import dataclasses
        cfg, C = self.cfg, self.tokens.shape[0]
        fi = RlcfEngine.tower_fwd_flops(self.visual)
        txt = dataclasses.replace(self.text, L=self.full_text_positions)
        ft = RlcfEngine.tower_fwd_flops(txt)
        total = cfg.n_views * fi + fi + C * ft + cfg.tta_steps * C * (ft + RlcfEngine.tower_dgrad_flops(txt))
        if self.scorer is not None and cfg.loss == "rlcf":
            total += cfg.n_selected * sum(RlcfEngine.tower_fwd_flops(t) for t in self.scorer.towers)
        return float(total)


class HostPipeline:
    """Overlaps the pinned-host -> device copy of batch i+1 with the adaptation of batch i.

    submit(images_pinned, slot) enqueues the H2D copy into staging buffer `slot` on a side stream;
    run(slot, out_pinned) makes the compute stream wait for that copy, moves the batch into the graph's static
    input (a device-to-device copy, ~0.1 ms per 300 MB), replays the captured step and copies the adapted logits
    back to pinned host memory.  Every byte of every batch still crosses PCIe inside the caller's timed region."""

    def __init__(self, eng: "RlcfEngine"):
        self.eng = eng
        self.copy_stream = torch.cuda.Stream()
        self.stage = [torch.empty_like(eng._static_images) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        cur = torch.cuda.current_stream()
        for e in self.free:
            e.record(cur)

    def submit(self, images_pinned: torch.Tensor, slot: int):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            self.stage[slot].copy_(images_pinned, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def run(self, slot: int, out_pinned: torch.Tensor) -> torch.Tensor:
        eng = self.eng
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready[slot])
        eng._static_images.copy_(self.stage[slot], non_blocking=True)
        self.free[slot].record(cur)
        eng._graph.replay()
        out_pinned.copy_(eng.logits_final, non_blocking=True)
        return out_pinned


class TextRunnerF32:
    """The text transformer in fp32 on the CUDA cores (rlcf_gemm_f32 / rlcf_attention_f32): the once-per-dataset class
    features keep the reference's precision (fp32 on the CPU; measured 2e-6 against it, where the fp16 tensor-core tower
    is at 8e-4 -- the text tower amplifies operand rounding ~3x more than the image tower)."""

    def __init__(self, w: TowerWeights, max_seq: int):
        if w.kind != "text" or w.layers[0].w32 is None:
            raise RlcfError("the fp32 text path needs a text tower prepared with keep_f32=True")
        self.w, self.max_seq = w, max_seq
        rows, d = max_seq * w.L, w.d
        f32 = dict(dtype=torch.float32, device=w.ln_flat.device)
        self.x = torch.empty(rows, d, **f32)
        self.a = torch.empty(rows, d, **f32)
        self.att = torch.empty(rows, d, **f32)
        self.qkv = torch.empty(rows, 3 * d, **f32)
        self.h = torch.empty(rows, 4 * d, **f32)

    def layers(self, n_seq: int) -> torch.Tensor:
        """Runs every residual block on self.x[:n_seq*L] in place (model.py:189-192, causal mask 328-334)."""
        w = self.w
        d, L, rows = w.d, w.L, n_seq * w.L
        ln = w.ln_flat
        x = self.x
        for l, lw in enumerate(w.layers):
            o1, o2 = w.ln_off("ln_1", l), w.ln_off("ln_2", l)
            ops.layernorm_fwd(x, ln[o1:], ln[o1 + d:], rows, d, out32=self.a)
            ops.gemm_f32(self.a, lw.w32["wqkv"], self.qkv, M=rows, bias=lw.bqkv)
            ops.attention_f32(self.qkv, n_seq, L, w.heads, self.att, causal=True)
            ops.gemm_f32(self.att, lw.w32["wo"], x, M=rows, epilogue=2, bias=lw.bo, resid=x)
            ops.layernorm_fwd(x, ln[o2:], ln[o2 + d:], rows, d, out32=self.a)
            ops.gemm_f32(self.a, lw.w32["wfc"], self.h, M=rows, epilogue=1, bias=lw.bfc)
            ops.gemm_f32(self.h, lw.w32["wproj"], x, M=rows, epilogue=2, bias=lw.bproj, resid=x)
        return x


def text_features(w: TowerWeights, tokens: torch.Tensor, chunk: int = 256, normalized: bool = True,
                  precise: bool | None = None) -> torch.Tensor:
    """L2-normalised text features [n, E] of tokenised prompts [n, ctx] (CLIP.encode_text, TPT/clip/model.py:342-356,
    followed by the normalisation of custom_clip.py:404-408 / clip_reward.py:139-150).  normalized=False returns the
    raw encode_text output.  precise (default: whenever the tower kept its fp32 weights) runs the transformer in fp32
    (TextRunnerF32) instead of on the fp16 tensor-core kernels: class features are computed once per dataset and feed
    every per-image step."""
    if w.kind != "text":
        raise RlcfError("text_features needs a text tower")
    n, L = tokens.shape
    if L != w.L:
        raise RlcfError(f"context length {L} != {w.L}")
    tokens = tokens.to(device=w.ln_flat.device, dtype=torch.int64).contiguous()
    if precise is None:
        precise = w.layers[0].w32 is not None and w.layers[0].wqkv.dim() == 2
    if precise:
        chunk = min(chunk, 128)
        frun = TextRunnerF32(w, min(chunk, n))
    run = TowerRunner(w, min(chunk, n)) if not precise else None
    out = torch.empty(n, w.E, dtype=torch.float32, device=w.ln_flat.device)
    inv = torch.empty(n, dtype=torch.float32, device=w.ln_flat.device)
    eot = tokens.argmax(dim=-1).to(torch.int32)   # eot_token is the highest id in each sequence (model.py:352-354)
    off = w.ln_off("ln_final")
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        tk = tokens[s:e].contiguous()
        rows = (torch.arange(e - s, device=tk.device, dtype=torch.int32) * L + eot[s:e]).contiguous()
        if precise:
            ops.embed_text(tk, w.tok_emb, w.pos, frun.x)
            x = frun.layers(e - s)
            ops.head_fwd(x, w.ln_flat[off:], w.ln_flat[off + w.d:], w.proj, e - s, w.d, w.E, feat=out[s:e],
                         inv_norm=inv[s:e], row_idx=rows, row_stride=L)
        else:
            x = run.forward(e - s, w.ln_flat, tokens=tk)
            run.head(x, e - s, w.ln_flat, row_idx=rows, feat=out[s:e], inv_norm=inv[s:e])
    return out if normalized else out / inv[:, None]


def image_features(w: TowerWeights, images: torch.Tensor, chunk: int = 256) -> torch.Tensor:
    """L2-normalised image features [n, E] (VisionTransformer.forward + normalisation, model.py:223-240)."""
    n = images.shape[0]
    images = images.float().contiguous()
    run = TowerRunner(w, min(chunk, n))
    out = torch.empty(n, w.E, dtype=torch.float32, device=w.ln_flat.device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = run.forward(e - s, w.ln_flat, images=images[s:e])
        run.head(x, e - s, w.ln_flat, feat=out[s:e])
    return out
