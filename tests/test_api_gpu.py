"""The reference-facing Python surface (clip.load / CLIPCLS_TTA / get_reward_model / test_time_tuning /
test_time_adapt_eval) driven exactly like TPT/tune_cls_rl.py drives the reference, checked against the oracle."""
import argparse
from copy import deepcopy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import parity_log as PL  # noqa: E402
from oracle import rlcf_oracle as O  # noqa: E402
from rlcf_b200 import clip, synthetic  # noqa: E402
from rlcf_b200.clip.custom_clip import CLIPCLS_TTA  # noqa: E402
from rlcf_b200.clip_reward import get_clip_reward, get_reward_model  # noqa: E402
from rlcf_b200 import tpt_cls_rl, tune_cls_rl  # noqa: E402  (not `from ... import test_*`: pytest would collect them)
from rlcf_b200.tpt_cls_rl import avg_entropy, select_confident_samples  # noqa: E402

DEV = torch.device("cuda:0")


def make_args(**kw):
    base = dict(tta_steps=1, selection_p=0.25, batch_size=16, min_entropy_reg=0, multiple_reward_models=0,
                reward_arch="synthetic:tiny-B:1", reward_amplify=0, sample_k=3, reward_process=1, process_batch=0,
                momentum_update=0, images_per_step=3, print_freq=1000, gpu=0)
    base.update(kw)
    return argparse.Namespace(**base)


def test_synthetic_weights_equal_oracle_generator():
    a, b = synthetic.make_state_dict("tiny-A", 3), O.make_clip_state_dict("tiny-A", 3)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)


def test_load_and_encode():
    model, embed_dim, preprocess = clip.load("synthetic:tiny-A:0", device=DEV)
    assert embed_dim == 128 and callable(preprocess)
    sd = O.make_clip_state_dict("tiny-A", 0)
    assert set(model.state_dict().keys()) == set(sd.keys())
    img = O.make_views(1, 4, 64, 2)
    tok = O.make_tokens(5, 512)
    with torch.no_grad():
        ref_i, ref_t = O.encode_image(sd, img), O.encode_text(sd, tok)
    got_i, got_t = model.encode_image(img.to(DEV)).cpu(), model.encode_text(tok.to(DEV)).cpu()
    assert (got_i - ref_i).abs().max() < 2e-3 * ref_i.abs().max()
    assert (got_t - ref_t).abs().max() < 3e-3 * ref_t.abs().max()
    with pytest.raises(RuntimeError):
        clip.load("ViT-Z/99", device=DEV)
    with pytest.raises(RuntimeError):
        clip.tokenize("word " * 100)
    assert clip.tokenize("word " * 100, truncate=True).shape == (1, 77)


def test_selection_and_entropy_helpers():
    torch.manual_seed(0)
    logits = torch.randn(16, 10, device=DEV) * 3
    out, idx = select_confident_samples(logits, 0.25)
    ref_out, ref_idx, _ = O.select_confident_samples(logits.cpu(), 0.25)
    assert torch.equal(idx.cpu(), ref_idx) and torch.equal(out.cpu(), ref_out)
    assert abs(avg_entropy(out).item() - O.avg_entropy(ref_out).item()) < 1e-5


def _setup(args, n_cls=10):
    tok_p, tok_r = O.make_tokens(n_cls, 512), O.make_tokens(n_cls, 512)
    model = CLIPCLS_TTA(DEV, [f"class {i}" for i in range(n_cls)], arch="synthetic:tiny-A:0",
                        prompt_prefix="a photo of a", only_norm=True, tokenized_prompts=tok_p)
    model = model.cuda(0)
    optimizer = torch.optim.AdamW(model.parameters(), 5e-3, weight_decay=5e-4)
    optim_state = deepcopy(optimizer.state_dict())
    reward_model = get_reward_model(DEV, args)
    reward_model.set_class_features(tokenized_classes=tok_r.to(DEV))
    scaler = torch.cuda.amp.GradScaler(init_scale=1000)
    sd_p, sd_r = O.make_clip_state_dict("tiny-A", 0), O.make_clip_state_dict("tiny-B", 1)
    return model, optimizer, optim_state, reward_model, scaler, sd_p, sd_r, tok_p, tok_r


def test_reference_style_loop_matches_oracle():
    assert get_clip_reward is get_reward_model
    args = make_args()
    model, optimizer, optim_state, reward_model, scaler, sd_p, sd_r, tok_p, tok_r = _setup(args)
    assert len(list(model.parameters())) == 2 * (2 * 2 + 2)       # LayerNorm weights and biases only
    cf, rc = O.class_features(sd_p, tok_p), O.class_features(sd_r, tok_r)
    assert (model.class_features.cpu() - cf).abs().max() < 3e-3
    views = O.make_views(2, 16, 64, 9)
    ocfg = O.OracleConfig(n_views=16, selection_p=0.25, tta_steps=1, sample_k=3, lr=5e-3)
    base = O.flat_ln_params(sd_p)
    for i in range(2):
        images = views[i * 16:(i + 1) * 16].to(DEV)
        model.reset()                                         # tune_cls_rl.py:210
        optimizer.load_state_dict(optim_state)                # tune_cls_rl.py:213
        assert torch.equal(model.clip_model.visual.ln_flat().cpu(), base)
        model.train()
        tpt_cls_rl.test_time_tuning(model, images, optimizer, scaler, args, reward_model=reward_model)
        model.eval()
        out = model(images[:1]).cpu()
        # the oracle sees the class features the CUDA text tower produced (inputs of the per-image loop)
        ref = O.adapt_one_image(sd_p, model.class_features.cpu(), views[i * 16:(i + 1) * 16], ocfg, sd_r,
                                reward_model.class_features.cpu())
        scale = ref["logits_all"].abs().max()
        delta = (ref["logits_final"] - ref["logits_all"][:1]).abs().max()
        PL.check_final_logits(f"api/ln_tuning/img{i}", out.numpy(), ref["logits_final"].numpy(), scale, delta,
                              allow_delta=0.02, why="sign-like AdamW steps on noise-level gradients (tests/test_parity_gpu.py GOLDEN_ALLOW); tiny towers, lr 5e-3")
        moved = (model.clip_model.visual.ln_flat().cpu() - base).abs()
        assert moved.max() > 1e-3                              # parameters really changed in place ...
        named = dict(model.clip_model.visual.named_parameters())
        assert (named["ln_post.weight"].detach().cpu() - sd_p["visual.ln_post.weight"]).abs().max() > 1e-4
        d = (model.clip_model.visual.ln_flat().cpu() - ref["params"]).abs()
        assert (d <= 0.02 * 5e-3).float().mean() > 0.9 and d.max() <= 2.02 * 5e-3
    model.reset()
    assert torch.equal(model.clip_model.visual.ln_flat().cpu(), base)   # ... and reset() restores them


def test_batched_eval_driver_equals_one_at_a_time():
    args = make_args()
    model, optimizer, optim_state, reward_model, scaler, *_ = _setup(args)
    views = O.make_views(7, 16, 64, 13)
    labels = torch.arange(7) % 10
    loader = [(views[i * 16:(i + 1) * 16], labels[i:i + 1]) for i in range(7)]   # ragged: 7 = 3 + 3 + 1
    res_batched = tune_cls_rl.test_time_adapt_eval(loader, model, optimizer, optim_state, scaler, args, device=DEV,
                                       reward_model=reward_model)
    args1 = make_args(images_per_step=1)
    res_single = tune_cls_rl.test_time_adapt_eval(loader, model, optimizer, optim_state, scaler, args1, device=DEV,
                                      reward_model=reward_model)
    assert res_batched == res_single
    assert 0.0 <= res_batched[0] <= res_batched[1] <= 100.0


def test_full_tuning_through_the_api():
    """rlcf-tune.sh's default mode (--tune_norm 0): every visual parameter adapts; reset() restores them."""
    args = make_args()
    tok = O.make_tokens(10, 512)
    model = CLIPCLS_TTA(DEV, [f"class {i}" for i in range(10)], arch="synthetic:tiny-A:0",
                        prompt_prefix="a photo of a", only_norm=False, tokenized_prompts=tok).cuda(0)
    optimizer = torch.optim.AdamW(model.parameters(), 1e-4, weight_decay=5e-4)
    optim_state = deepcopy(optimizer.state_dict())
    assert len(list(model.parameters())) == len(list(model.clip_model.visual.parameters())) > 12
    reward_model = get_reward_model(DEV, args)
    reward_model.set_class_features(tokenized_classes=tok.to(DEV))
    sd_p, sd_r = O.make_clip_state_dict("tiny-A", 0), O.make_clip_state_dict("tiny-B", 1)
    views = O.make_views(1, 16, 64, 23)
    model.reset()
    optimizer.load_state_dict(optim_state)
    model.train()
    tpt_cls_rl.test_time_tuning(model, views.to(DEV), optimizer, None, args, reward_model=reward_model)
    model.eval()
    out = model(views[:1].to(DEV)).cpu()
    ocfg = O.OracleConfig(n_views=16, selection_p=0.25, tta_steps=1, sample_k=3, lr=1e-4)
    ref = O.adapt_one_image(sd_p, model.class_features.cpu(), views, ocfg, sd_r, reward_model.class_features.cpu(),
                            tune="full")
    scale = ref["logits_all"].abs().max()
    delta = (ref["logits_final"] - ref["logits_all"][:1]).abs().max()
    PL.check_final_logits("api/full_tuning", out.numpy(), ref["logits_final"].numpy(), scale, delta, allow_delta=0.02,
                          why="sign-like AdamW steps on noise-level gradients (tests/test_parity_gpu.py GOLDEN_ALLOW); tiny towers, lr 5e-3")
    w = dict(model.clip_model.visual.named_parameters())["transformer.resblocks.0.mlp.c_fc.weight"].detach().cpu()
    assert (w - sd_p["visual.transformer.resblocks.0.mlp.c_fc.weight"]).abs().max() > 5e-5     # weights really moved
    model.reset()
    w = dict(model.clip_model.visual.named_parameters())["transformer.resblocks.0.mlp.c_fc.weight"].detach().cpu()
    assert torch.equal(w, sd_p["visual.transformer.resblocks.0.mlp.c_fc.weight"])


def test_unsupported_modes_fail_loudly():
    args = make_args()
    with pytest.raises(NotImplementedError):
        get_reward_model(DEV, make_args(multiple_reward_models=1))
    with pytest.raises(Exception):
        clip.load("synthetic:tiny-A:0", device="cpu")[0].encode_image(torch.zeros(1, 3, 64, 64))  # no CPU fallback


def test_prompt_tuning_api_matches_oracle():
    """get_coop / ClipTestTimeTuning / PromptLearner driven like TPT/tpt_cls_rl.py:219-279."""
    from rlcf_b200.clip.custom_clip import get_coop
    # reward seed 6: every CLIPScore of the sampled classes is positive (with seed 1 most are clipped to 0 and the
    # rewards -- hence the whole adaptation -- vanish: the round-1 form of this test compared two un-adapted models)
    args = make_args(reward_arch="synthetic:tiny-Q:6", tta_steps=1)
    tokens = O.make_tokens(9, 49408)
    tokens[:, 1:5] = torch.tensor([320, 1125, 539, 320])     # every prompt starts with the 4 context tokens
    model = get_coop("synthetic:tiny-P:0", "synthetic", DEV, 4, None, classnames=[f"c{i}" for i in range(9)],
                     tokenized_prompts=tokens)
    sd_p, sd_r = O.make_clip_state_dict("tiny-P", 0), O.make_clip_state_dict("tiny-Q", 6)
    with torch.no_grad():   # random ctx init in the model -> use it as the oracle's starting point
        ctx_init = model.prompt_learner.ctx.detach().cpu().clone()
    for n, p in model.named_parameters():
        if "prompt_learner" not in n:
            p.requires_grad_(False)
    optimizer = torch.optim.AdamW(model.prompt_learner.parameters(), 5e-3, weight_decay=5e-4)
    optim_state = deepcopy(optimizer.state_dict())
    reward_model = get_reward_model(DEV, args)
    reward_model.set_class_features(tokenized_classes=model.prompt_learner.tokenized_prompts)
    views = O.make_views(1, 16, 64, 17)
    # plain inference through the API == oracle forward with the same context
    with torch.no_grad():
        ref0 = O.policy_logits.__globals__["encode_image"](sd_p, views[:2])
        ref0 = ref0 / ref0.norm(dim=-1, keepdim=True)
        ref_logits = 100.0 * ref0 @ O.prompt_text_features(sd_p, tokens, ctx_init).t()
    got = model(views[:2].to(DEV)).cpu()
    assert (got - ref_logits).abs().max() <= 2.5e-3 * ref_logits.abs().max()
    model.reset()
    optimizer.load_state_dict(optim_state)
    tpt_cls_rl.test_time_tuning(model, views.to(DEV), optimizer, None, args, reward_model=reward_model)
    out = model(views[:1].to(DEV)).cpu()
    ocfg = O.OracleConfig(n_views=16, selection_p=0.25, tta_steps=1, sample_k=3, lr=5e-3)
    ref = O.adapt_one_image_prompt(sd_p, tokens, ctx_init, views, ocfg, sd_r, reward_model.class_features.cpu())
    scale = ref["logits_all"].abs().max()
    delta = (ref["logits_final"][0] - ref["logits_all"][0]).abs().max()
    PL.check_final_logits("api/prompt_tuning", out[0].numpy(), ref["logits_final"][0].numpy(), scale, delta,
                          allow_delta=0.02, why="sign-like AdamW steps on noise-level gradients (tests/test_parity_gpu.py GOLDEN_ALLOW); tiny towers, lr 5e-3")
    d = (model.prompt_learner.ctx.detach().cpu().flatten() - ref["params"]).abs()
    assert d.max() <= 2.02 * 5e-3 and (d <= 0.02 * 5e-3).float().mean() > 0.9
    model.reset()
    assert torch.equal(model.prompt_learner.ctx.detach().cpu(), ctx_init)


@pytest.mark.parametrize("name", ["tiny_prompt_middle", "tiny_prompt_learned_cls"])
def test_prompt_layouts_through_the_api(name):
    """ClipTestTimeTuning(ctx_position="middle") and (learned_cls=True) through the reference-style loop
    (tpt_cls_rl.py:219-279): PromptLearner builds the layout itself (real BPE tokens), test_time_tuning adapts the
    context (and class) vectors in place; against the reference's own outputs (tests/golden, oracle/make_golden.py)."""
    import ast
    import os
    from rlcf_b200.clip import simple_tokenizer as ST
    from rlcf_b200.clip.custom_clip import ClipTestTimeTuning
    if ST._find_vocab() is None:
        pytest.skip("OpenAI BPE vocabulary not available on this machine")
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    cfg = ast.literal_eval(str(z["meta"]))
    names = ["tench", "goldfish", "great white shark", "tiger shark", "hammerhead", "electric ray", "stingray",
             "cock", "hen", "ostrich", "brambling", "goldfinch"][:cfg["C"]]
    learned = bool(cfg.get("learned_cls", False))
    model = ClipTestTimeTuning(DEV, names, None, arch="synthetic:" + cfg["policy"] + ":0", n_ctx=4,
                               ctx_init=cfg["ctx_init"], ctx_position=cfg.get("ctx_position", "end"), learned_cls=learned,
                               cls_init=torch.tensor(z["cls_init"]) if learned else None)
    assert np.array_equal(model.prompt_learner.tokenized_prompts.cpu().numpy(), z["tokens"])
    for n, p in model.named_parameters():
        if "prompt_learner" not in n:
            p.requires_grad_(False)
    optimizer = torch.optim.AdamW(model.prompt_learner.parameters(), cfg["lr"], weight_decay=5e-4)
    optim_state = deepcopy(optimizer.state_dict())
    reward_model = None
    if cfg["loss"] == "rlcf":
        args = make_args(reward_arch=f"synthetic:{cfg['reward']}:{cfg['reward_seed']}", tta_steps=cfg["steps"])
        reward_model = get_reward_model(DEV, args)
        reward_model.set_class_features(tokenized_classes=model.prompt_learner.tokenized_prompts)
    else:
        args = make_args(tta_steps=cfg["steps"])
    args.selection_p, args.batch_size = cfg["rho"], cfg["V"]
    views = O.make_views(cfg["n_img"], cfg["V"], 64, cfg["view_seed"])[:cfg["V"]]
    model.reset()
    optimizer.load_state_dict(optim_state)
    tpt_cls_rl.test_time_tuning(model, views.to(DEV), optimizer, None, args, reward_model=reward_model)
    out = model(views[:1].to(DEV)).cpu().numpy()[0]
    scale = np.abs(z["img0.logits_all"]).max()
    delta = np.abs(z["img0.logits_final"][0] - z["img0.logits_all"][0]).max()
    PL.check_final_logits(f"api/{name}", out, z["img0.logits_final"][0], scale, delta, allow_delta=0.02,
                          why="prompt tuning on 128-wide towers (tests/test_parity_gpu.py PROMPT_ALLOW['tiny'])")
    assert out.argmax() == z["img0.logits_final"][0].argmax()
    d = (model.prompt_learner.learnable_flat().cpu() - torch.tensor(z["img0.params"])).abs()
    assert d.max() <= 2.02 * cfg["lr"] * cfg["steps"] + 1e-7
    assert (d <= 0.05 * cfg["lr"] * cfg["steps"]).float().mean() >= 0.9
    model.reset()
    assert torch.equal(model.prompt_learner.ctx.detach().cpu(), torch.tensor(z["ctx_init"]))


def test_reward_model_ensemble_through_the_api():
    """CLIPRewardsMultiple (TPT/clip_reward.py:180-307) with two ViT members behind get_reward_model's
    --multiple_reward_models switch: CLIPScore as the reference computes it, and the adapted logits of the reference-
    style loop against the oracle's ensemble restatement (itself pinned to the reference's class by golden vectors)."""
    from rlcf_b200 import clip_reward as CR
    args = make_args(multiple_reward_models=1, reward_arch="synthetic:tiny-B:1,synthetic:tiny-A:5")
    with pytest.raises(NotImplementedError):
        CR.get_reward_model(DEV, make_args(multiple_reward_models=1))          # the reference's RN50x64 ensemble
    tok_p, tok_r = O.make_tokens(10, 512), O.make_tokens(10, 512)
    model = CLIPCLS_TTA(DEV, [f"class {i}" for i in range(10)], arch="synthetic:tiny-A:0", prompt_prefix="a photo of a",
                        only_norm=True, tokenized_prompts=tok_p).cuda(0)
    optimizer = torch.optim.AdamW(model.parameters(), 5e-3, weight_decay=5e-4)
    optim_state = deepcopy(optimizer.state_dict())
    reward_model = CR.CLIPRewardsMultiple(DEV, arch=["synthetic:tiny-B:1", "synthetic:tiny-A:5"], sample_k=3,
                                          process_batch=False, default_resolutions=64)
    assert reward_model.weights == [0.5, 0.5] and reward_model.n_model == 2
    reward_model.set_class_features(tokenized_classes=tok_r.to(DEV))
    sd_p = O.make_clip_state_dict("tiny-A", 0)
    sd_r = [O.make_clip_state_dict("tiny-B", 1), O.make_clip_state_dict("tiny-A", 5)]
    views = O.make_views(1, 16, 64, 9)
    images = views.to(DEV)
    # CLIPScore of the class: weighted sum of the members' clipped cosines
    reward_model.set_image_features(images[:2])
    idx = torch.tensor([1, 4, 7, 0, 2, 3], device=DEV)
    got = reward_model.CLIPScore(class_index=idx, pairwise=False).cpu()
    want = O.clip_score_multi([f.cpu() for f in reward_model.class_features], [f.cpu() for f in reward_model.image_features],
                              idx.cpu(), 3, reward_model.weights)
    assert torch.allclose(got, want, atol=1e-6)
    # the loop
    model.reset()
    optimizer.load_state_dict(optim_state)
    model.train()
    tpt_cls_rl.test_time_tuning(model, images, optimizer, torch.cuda.amp.GradScaler(init_scale=1000), args,
                                reward_model=reward_model)
    model.eval()
    out = model(images[:1]).cpu()
    ocfg = O.OracleConfig(n_views=16, selection_p=0.25, tta_steps=1, sample_k=3, lr=5e-3, reward_weights=(0.5, 0.5))
    ref = O.adapt_one_image(sd_p, model.class_features.cpu(), views, ocfg, sd_r,
                            [f.cpu() for f in reward_model.class_features])
    scale = ref["logits_all"].abs().max()
    delta = (ref["logits_final"] - ref["logits_all"][:1]).abs().max()
    PL.check_final_logits("api/multi_reward", out.numpy(), ref["logits_final"].numpy(), scale, delta, allow_delta=0.02,
                          why="sign-like AdamW steps on noise-level gradients (tests/test_parity_gpu.py GOLDEN_ALLOW); tiny towers, lr 5e-3")
    assert isinstance(reward_model.image_features, list) and len(reward_model.image_features) == 2


def test_eval_driver_rebuilds_class_prompts_per_test_set(tmp_path):
    """ADVICE r1 (medium), on the device: the ImageFolder driver gives every test set its own label space -- class
    names from a --classnames table where the folders are WordNet ids, from the folder names otherwise -- by calling
    model.reset_classnames_and_state + reward_model.set_class_features per set (tune_cls_rl.py:120-143), generates the
    views on the GPU, adapts a ragged stream (6 images, 4 per launch sequence) and refuses id folders without a table."""
    import json
    import os
    from PIL import Image
    from rlcf_b200.clip import simple_tokenizer as ST
    from rlcf_b200.params import get_args
    if ST._find_vocab() is None:
        pytest.skip("OpenAI BPE vocabulary not available on this machine")
    rng = np.random.default_rng(0)
    root = tmp_path / "data"
    layout = {"imagenet-a": ["n01440764", "n01443537", "n01484850"], "oxford_pets": ["golden_retriever", "tabby_cat"]}
    for d, classes in layout.items():
        for c in classes:
            os.makedirs(root / d / c)
            for k in range(2 if d == "imagenet-a" else 3):
                arr = rng.integers(0, 255, size=(96 + 8 * k, 120, 3), dtype=np.uint8)
                Image.fromarray(arr).save(root / d / c / f"{k}.png")
    table = tmp_path / "names.json"
    table.write_text(json.dumps({"n01440764": "tench", "n01443537": "goldfish", "n01484850": "great white shark"}))
    argv = [str(root), "--test_sets", "A/pets", "-a", "ViT-B/32", "--reward_arch", "ViT-B/32", "--tpt", "--tune_norm", "1",
            "--batch_size", "8", "--selection_p", "0.5", "--tta_steps", "1", "--sample_k", "2", "--synthetic_weights",
            "--images_per_step", "4", "--workers", "0", "--output", str(tmp_path / "out"), "--ctx_init", "a_photo_of_a"]
    seen = []
    from rlcf_b200.clip import custom_clip
    orig = custom_clip.CLIPCLS_TTA.reset_classnames_and_state

    def spy(self, classnames, arch, tokenized_prompts=None):
        seen.append(list(classnames))
        return orig(self, classnames, arch, tokenized_prompts)

    custom_clip.CLIPCLS_TTA.reset_classnames_and_state = spy
    try:
        results = tune_cls_rl.main_worker(0, get_args(argv + ["--classnames", str(table)]))
    finally:
        custom_clip.CLIPCLS_TTA.reset_classnames_and_state = orig
    assert set(results) == {"A", "pets"} and all(len(v) == 2 for v in results.values())
    assert seen == [["tench", "goldfish", "great white shark"], ["golden retriever", "tabby cat"]]
    with pytest.raises(SystemExit):
        tune_cls_rl.main_worker(0, get_args(argv))          # wnid folders, no table: never "a photo of a n01440764."
