"""The three "only work that can reach the output" restructurings of round 2, shown on the CPU
against the oracle's restatement of the reference in fp32 -- i.e. as properties of the reference's algorithm, independent
of any kernel:
  * the last block of VisionTransformer.forward only matters through its class-token row (model.py:232-238);
  * the causal text tower's EOT row does not depend on the padding behind the EOT (model.py:328-334, 352-354);
  * the gradient w.r.t. the context vectors is the same on the truncated sequence.
The GPU tests (test_class_token_only_last_block_equals_the_full_forward,
test_text_tower_on_the_eot_prefix_equals_all_77_positions, test_adopted_activations_equal_a_second_forward) check the
CUDA path against itself; these check the idea against the reference's maths."""
import torch
import torch.nn.functional as F

from oracle import rlcf_oracle as O


def _class_token_only_encode_image(sd, images):
    """encode_image with the last block evaluated for the class-token row only: ln_1 and the K / V projections for every
    token, everything else of that block for row 0 (what TowerRunner.forward does on the GPU)."""
    w = sd["visual.conv1.weight"]
    d, p = w.shape[0], w.shape[-1]
    heads, hd = d // 64, 64
    x = F.conv2d(images, w, stride=p)
    x = x.reshape(x.shape[0], d, -1).permute(0, 2, 1)
    x = torch.cat([sd["visual.class_embedding"].expand(x.shape[0], 1, d), x], dim=1) + sd["visual.positional_embedding"]
    x = F.layer_norm(x, (d,), sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"], 1e-5)
    n = O._n_layers(sd, "visual.")
    for l in range(n - 1):
        x = O._block(x, sd, f"visual.transformer.resblocks.{l}.", heads, None)
    rb = f"visual.transformer.resblocks.{n - 1}."
    N, L, _ = x.shape
    h = F.layer_norm(x, (d,), sd[rb + "ln_1.weight"], sd[rb + "ln_1.bias"], 1e-5)
    wq, wkv = sd[rb + "attn.in_proj_weight"][:d], sd[rb + "attn.in_proj_weight"][d:]
    bq, bkv = sd[rb + "attn.in_proj_bias"][:d], sd[rb + "attn.in_proj_bias"][d:]
    q = (h[:, 0] @ wq.t() + bq).view(N, heads, 1, hd)                        # the class token's query only
    k, v = (h @ wkv.t() + bkv).view(N, L, 2, heads, hd).permute(2, 0, 3, 1, 4)
    a = (((q * hd ** -0.5) @ k.transpose(-1, -2)).softmax(-1) @ v).reshape(N, d)
    c = x[:, 0] + a @ sd[rb + "attn.out_proj.weight"].t() + sd[rb + "attn.out_proj.bias"]
    u = F.layer_norm(c, (d,), sd[rb + "ln_2.weight"], sd[rb + "ln_2.bias"], 1e-5) @ sd[rb + "mlp.c_fc.weight"].t() \
        + sd[rb + "mlp.c_fc.bias"]
    c = c + (u * torch.sigmoid(1.702 * u)) @ sd[rb + "mlp.c_proj.weight"].t() + sd[rb + "mlp.c_proj.bias"]
    return F.layer_norm(c, (d,), sd["visual.ln_post.weight"], sd["visual.ln_post.bias"], 1e-5) @ sd["visual.proj"]


def test_last_block_only_matters_through_the_class_token():
    for arch, seed in (("tiny-A", 0), ("tiny-B", 1)):
        sd = O.make_clip_state_dict(arch, seed)
        img = O.make_views(1, 4, O.ARCHS[arch][1], 5)
        with torch.no_grad():
            ref, got = O.encode_image(sd, img), _class_token_only_encode_image(sd, img)
        assert (ref - got).abs().max() <= 2e-6 * ref.abs().max()


def _prefix_state(sd, n):
    return dict(sd, positional_embedding=sd["positional_embedding"][:n])


def test_eot_row_of_the_causal_text_tower_ignores_the_padding():
    sd = O.make_clip_state_dict("tiny-A", 0)
    tok = O.make_tokens(9, O.ARCHS["tiny-A"][6], seed=3)
    need = int(tok.argmax(-1).max()) + 1
    n = (need + 7) // 8 * 8
    assert n < tok.shape[1]
    with torch.no_grad():
        full = O.encode_text(sd, tok)
        pref = O.encode_text(_prefix_state(sd, n), tok[:, :n])
    assert (full - pref).abs().max() <= 2e-6 * full.abs().max()


def test_context_gradient_is_the_same_on_the_prefix():
    sd = O.make_clip_state_dict("tiny-A", 0)
    tok = O.make_tokens(6, O.ARCHS["tiny-A"][6], seed=4)
    n = (int(tok.argmax(-1).max()) + 1 + 7) // 8 * 8
    target = torch.randn(6, O.ARCHS["tiny-A"][0], generator=torch.Generator().manual_seed(1))
    grads = []
    for state, t in ((sd, tok), (_prefix_state(sd, n), tok[:, :n])):
        ctx = (sd["token_embedding.weight"][tok[0, 1:3]]).clone().requires_grad_(True)     # two learnable positions
        emb = sd["token_embedding.weight"][t].clone()
        emb = torch.cat([emb[:, :1], ctx.unsqueeze(0).expand(t.shape[0], -1, -1), emb[:, 3:]], dim=1)
        (O.text_from_embeddings(state, emb, t) * target).sum().backward()
        grads.append(ctx.grad.clone())
    assert (grads[0] - grads[1]).abs().max() <= 5e-6 * grads[0].abs().max()
