"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol of include/rlcf_b200.h, the host
logic that needs no device (tokenizer, flags, flat LayerNorm layout, sharding, FLOP accounting) behaves, and a
world_size-2 gloo run of the accuracy-counter reduction matches the single-process result."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from rlcf_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = _lib.header_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(handle, s)]
    assert not missing, missing
    assert set(_lib._PROTOS) == set(syms), set(_lib._PROTOS) ^ set(syms)
    handle.rlcf_abi_version.restype = ctypes.c_int
    assert handle.rlcf_abi_version() == 1


def test_ops_refuse_cpu_tensors():
    from rlcf_b200 import _lib, ops
    a = torch.zeros(8, 64, dtype=torch.float16)
    with pytest.raises(_lib.RlcfError):
        ops.gemm(a, a, torch.zeros(8, 8, dtype=torch.float16))


def test_flag_names_and_defaults_match_reference():
    from rlcf_b200.params import build_parser
    a = build_parser().parse_args([])
    # TPT/params.py:23-73
    assert (a.arch, a.batch_size, a.selection_p, a.tta_steps, a.lr, a.weight_decay) == ("RN50", 64, 0.1, 1, 5e-3, 5e-4)
    assert (a.sample_k, a.reward_arch, a.reward_process, a.process_batch, a.reward_amplify) == (5, "ViT-L/14", 1, 0, 0)
    assert (a.multiple_reward_models, a.tune_norm, a.momentum_update, a.ctx_init, a.n_ctx, a.seed) == (0, 0, 0, None, 4, 0)


def test_flat_layernorm_layout_and_flops():
    from rlcf_b200 import engine as E
    w = E.TowerWeights(kind="visual", d=768, heads=12, n_layers=12, L=197, E=512, patch=16)
    w.ln_flat = torch.zeros((4 * 12 + 4) * 768)
    assert w.P == 39936
    assert w.ln_off("ln_pre") == 0 and w.ln_off("ln_1", 0) == 1536 and w.ln_off("ln_2", 11) == 1536 + 11 * 3072 + 1536
    assert w.ln_off("ln_post") == 39936 - 1536
    names = [k for k, _ in w.ln_names("visual.")]
    assert names[0] == "visual.ln_pre.weight" and names[-1] == "visual.ln_post.bias" and len(names) == 52 * 2 // 2
    offs = [o for _, o in w.ln_names("visual.")]
    assert offs == sorted(offs) and len(set(offs)) == len(offs)
    assert abs(E.RlcfEngine.tower_fwd_flops(w) / 1e9 - 35.127) < 0.01
    assert abs(E.RlcfEngine.tower_dgrad_flops(w) / 1e9 - 36.33) < 0.01
    r = E.TowerWeights(kind="visual", d=1024, heads=16, n_layers=24, L=257, E=768, patch=14)
    assert abs(E.RlcfEngine.tower_fwd_flops(r) / 1e9 - 162.026) < 0.01
    cfg = E.RlcfConfig()
    assert cfg.n_selected == 6 and E.RlcfConfig(n_views=8).n_selected == 0


def _vocab_path():
    from rlcf_b200.clip import simple_tokenizer as ST
    return ST._find_vocab()


def test_tokenizer_matches_reference_vectors():
    """With OpenAI's merge table (shipped next to the tokenizer by build()), token ids equal the reference tokenizer's:
    known ids, and every prompt of the fixture that the reference tokenised itself (tests/golden/b32_cfg1_exact.npz,
    tiny_prompt_rlcf_2step.npz: `prompt_learner.tokenized_prompts`, custom_clip.py:140-150)."""
    import numpy as np
    vocab = _vocab_path()
    if vocab is None:
        pytest.skip("OpenAI BPE vocabulary not available on this machine")
    from rlcf_b200.clip import clip
    from rlcf_b200.clip.simple_tokenizer import SimpleTokenizer
    tok = SimpleTokenizer(vocab)
    assert not tok.byte_fallback
    assert tok.encode("a photo of a great white shark.") == [320, 1125, 539, 320, 830, 1579, 7980, 269]
    assert tok.encoder["<|startoftext|>"] == 49406 and tok.encoder["<|endoftext|>"] == 49407
    assert tok.decode(tok.encode("a photo of a tench.")).strip() == "a photo of a tench ."
    g = os.path.join(ROOT, "tests", "golden")
    z = np.load(os.path.join(g, "b32_cfg1_exact.npz"))
    mine = clip.tokenize([f"a photo of a class {i}." for i in range(32)])
    assert np.array_equal(mine.numpy(), z["tokens"])
    z = np.load(os.path.join(g, "tiny_prompt_rlcf_2step.npz"))
    names = ["tench", "goldfish", "great white shark", "tiger shark", "hammerhead", "electric ray", "stingray",
             "cock", "hen", "ostrich", "brambling", "goldfinch"]
    assert np.array_equal(clip.tokenize([f"a photo of a {n}." for n in names]).numpy(), z["tokens"])
    with pytest.raises(RuntimeError):
        clip.tokenize("word " * 100)                       # clip.py:230: too long for the context length
    assert clip.tokenize("word " * 100, truncate=True)[0, -1] == 49407


def test_missing_vocabulary_is_a_hard_error(monkeypatch):
    """No silent degraded path: without the merge table the tokenizer raises, unless the byte-level fallback was asked
    for explicitly (synthetic-weight runs)."""
    from rlcf_b200.clip import simple_tokenizer as ST
    monkeypatch.setattr(ST, "_find_vocab", lambda: None)
    monkeypatch.delenv("RLCF_BPE_FALLBACK", raising=False)
    with pytest.raises(RuntimeError, match="bpe_simple_vocab"):
        ST.SimpleTokenizer()
    tok = ST.SimpleTokenizer(allow_byte_fallback=True)
    assert tok.byte_fallback
    ids = tok.encode("a photo of a dog.")
    assert ids and max(ids) < 49406


def _script_archive(sd, path):
    """An OpenAI-style TorchScript archive: a scripted module whose hierarchy reproduces the state-dict keys, plus the
    three metadata buffers the real archives carry (clip.py:126-142 / model.py:430-432 drop them)."""
    import torch.nn as nn

    class Holder(nn.Module):
        def forward(self, x):
            return x

    root = Holder()
    for key, val in sd.items():
        mod, parts = root, key.split(".")
        for part in parts[:-1]:
            if not hasattr(mod, part):
                mod.add_module(part, Holder())
            mod = getattr(mod, part)
        mod.register_parameter(parts[-1], nn.Parameter(val.clone(), requires_grad=False))
    root.register_buffer("input_resolution", torch.tensor(64))
    root.register_buffer("context_length", torch.tensor(77))
    root.register_buffer("vocab_size", torch.tensor(512))
    torch.jit.save(torch.jit.script(root), path)


def test_clip_load_reads_torchscript_archives_and_state_dicts(tmp_path):
    """clip.load(path) (TPT/clip/clip.py:121-142): TorchScript archive -> state_dict -> build_model -> fp32 model,
    3-tuple result; a plain state_dict file and a {'state_dict': ...} checkpoint load the same weights."""
    from rlcf_b200 import synthetic
    from rlcf_b200.clip import clip
    sd = synthetic.make_state_dict("tiny-A", 4)
    jit_path, sd_path, ck_path = (str(tmp_path / n) for n in ("tiny.pt", "tiny_sd.pt", "tiny_ck.pt"))
    _script_archive(sd, jit_path)
    torch.save(sd, sd_path)
    torch.save({"state_dict": sd}, ck_path)
    assert set(torch.jit.load(jit_path).state_dict()) == set(sd) | {"input_resolution", "context_length", "vocab_size"}
    for path in (jit_path, sd_path, ck_path):
        model, embed_dim, preprocess = clip.load(path, device="cpu")
        assert embed_dim == 128 and model.visual.input_resolution == 64 and callable(preprocess)
        got = model.state_dict()
        assert set(got) == set(sd)
        assert all(torch.equal(got[k], sd[k]) and got[k].dtype == torch.float32 for k in sd)
    with pytest.raises(RuntimeError):
        clip.load("no-such-model", device="cpu")
    with pytest.raises(RuntimeError):
        clip.load("ViT-B/16", device="cpu", download_root=str(tmp_path))   # known name, file absent, no network


def test_synthetic_dataset_sharding():
    from rlcf_b200.tune_cls_rl import SyntheticViews
    full = SyntheticViews(10, 4, 32, 5, 0)
    shards = [SyntheticViews(10, 4, 32, 5, 0, r, 3) for r in range(3)]
    assert sorted(i for s in shards for i in s.idx) == full.idx
    v, y = shards[1][0]
    v0, y0 = full[1]
    assert torch.equal(v, v0) and int(y) == int(y0) == 1 and v.shape == (4, 3, 32, 32)


GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from rlcf_b200.tune_cls_rl import SyntheticViews
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
ds = SyntheticViews(11, 2, 8, 4, 5, r, w)
# stand-in for the adapted prediction: a deterministic function of the sample, so shards can be compared
hits = torch.zeros(3, dtype=torch.int64)
for i in range(len(ds)):
    v, y = ds[i]
    pred = int(v.sum().item() * 1000) % 4
    hits += torch.tensor([int(pred == int(y)), 1, 1])
dist.all_reduce(hits, op=dist.ReduceOp.SUM)
if r == 0:
    print("HITS", hits.tolist())
dist.destroy_process_group()
"""


def test_two_rank_gloo_counter_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("HITS")][0]
    from rlcf_b200.tune_cls_rl import SyntheticViews
    ds = SyntheticViews(11, 2, 8, 4, 5)
    exp = [0, 0, 0]
    for i in range(len(ds)):
        v, y = ds[i]
        exp[0] += int(int(v.sum().item() * 1000) % 4 == int(y))
        exp[1] += 1
        exp[2] += 1
    assert line == f"HITS {exp}"


def test_ctypes_prototypes_match_the_header_parameter_counts():
    """Every entry point's ctypes argtypes list has exactly as many entries as the C declaration has parameters
    (guards against the binding drifting from include/rlcf_b200.h when a signature changes)."""
    import re
    from rlcf_b200 import _lib
    with open(_lib.HEADER_PATH) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    decls = re.findall(r"\b(rlcf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(decls) >= 40
    for name, params in decls:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert len(_lib._PROTOS[name]) == n, f"{name}: header has {n} parameters, _lib._PROTOS {len(_lib._PROTOS[name])}"


def _stub_cls_tta(only_norm, momentum_update, update_freq=2, update_w=0.5, momentum=0.9):
    """CLIPCLS_TTA without its constructor (which loads a CLIP onto the GPU): only the weight-snapshot logic."""
    import copy
    import types

    from rlcf_b200.clip.custom_clip import CLIPCLS_TTA

    class Visual(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ln_pre = torch.nn.LayerNorm(8)
            self.fc = torch.nn.Linear(8, 8)

    m = object.__new__(CLIPCLS_TTA)
    torch.nn.Module.__init__(m)
    torch.manual_seed(0)
    m.clip_model = types.SimpleNamespace(visual=Visual())
    m.only_norm, m.momentum_update = only_norm, momentum_update
    m.update_freq, m.update_w, m.momentum, m.update_counter = update_freq, update_w, momentum, 0
    sd = m.clip_model.visual.state_dict()
    m.clip_state_dict, m.initial_state_dict = copy.deepcopy(sd), copy.deepcopy(sd)
    m.momentum_state_dict = copy.deepcopy(sd)
    return m


def test_reset_restores_what_the_mode_can_change():
    """reset() (TPT/clip/custom_clip.py:456-458): whole tower in full-tuning mode, LayerNorm entries in LN mode."""
    for only_norm in (False, True):
        m = _stub_cls_tta(only_norm, False)
        vis = m.clip_model.visual
        w0, g0 = vis.fc.weight.detach().clone(), vis.ln_pre.weight.detach().clone()
        with torch.no_grad():
            vis.ln_pre.weight.add_(1.0)
            if not only_norm:
                vis.fc.weight.add_(1.0)
        m.reset()
        assert torch.equal(vis.ln_pre.weight, g0) and torch.equal(vis.fc.weight, w0)
        assert len(list(m.parameters())) == (2 if only_norm else 4)


def test_momentum_update_model_follows_the_reference_formula():
    """custom_clip.py:460-475: EMA after every sample; every update_freq samples the reset state becomes
    (1 - update_w) * pretrained + update_w * EMA."""
    m = _stub_cls_tta(False, True, update_freq=2, update_w=0.5, momentum=0.9)
    vis = m.clip_model.visual
    w_pre = vis.fc.weight.detach().clone()
    ema = w_pre.clone()
    for step in range(1, 4):
        with torch.no_grad():
            vis.fc.weight.add_(0.1 * step)                 # "adapted" weights of this sample
        ema = 0.9 * ema + (1.0 - 0.9) * vis.fc.weight.detach()
        m.momentum_update_model()
        assert torch.allclose(m.momentum_state_dict["fc.weight"], ema, atol=0, rtol=0)
        if step == 2:                                      # counter reached update_freq
            assert m.update_counter == 0
            assert torch.equal(m.initial_state_dict["fc.weight"], (1 - 0.5) * w_pre + 0.5 * ema)
            init2 = m.initial_state_dict["fc.weight"].clone()
        elif step == 1:
            assert torch.equal(m.initial_state_dict["fc.weight"], w_pre)
        m.reset()                                          # next sample starts from the (possibly moved) reset state
        assert torch.equal(vis.fc.weight, m.initial_state_dict["fc.weight"])
    assert torch.equal(m.initial_state_dict["fc.weight"], init2) and m.update_counter == 1
    off = _stub_cls_tta(False, False)
    off.momentum_update_model()
    assert off.update_counter == 0


def test_class_names_per_test_set(tmp_path):
    """folder names that are ids never become prompts; every test set gets its own label space."""
    import json
    from rlcf_b200.tune_cls_rl import class_names_for, load_classname_table
    assert class_names_for(["golden_retriever", "tabby_cat"], None, "pets") == ["golden retriever", "tabby cat"]
    with pytest.raises(SystemExit):
        class_names_for(["n01440764", "n01443537"], None, "A")
    with pytest.raises(SystemExit):
        class_names_for(["0", "1", "10"], None, "V")
    p = tmp_path / "map.txt"
    p.write_text("n01440764 tench, Tinca tinca\nn01443537 goldfish, Carassius auratus\n")
    table = load_classname_table(str(p))
    assert class_names_for(["n01443537", "n01440764"], table, "A") == ["goldfish", "tench"]
    with pytest.raises(SystemExit):
        class_names_for(["n01443537", "n09999999"], table, "A")
    assert class_names_for(["tabby_cat", "pug"], table, "pets") == ["tabby cat", "pug"]      # named in clear: no entry needed
    p = tmp_path / "names.json"
    p.write_text(json.dumps({"V": ["zero", "one", "two"] + [f"c{i}" for i in range(3, 11)], "A": {"n01440764": "tench"}}))
    table = load_classname_table(str(p))
    assert class_names_for(["0", "1", "10", "2"], table, "V") == ["zero", "one", "c10", "two"]   # ImageFolder sorts as text
    assert class_names_for(["n01440764"], table, "A") == ["tench"]


def test_unsupported_reference_flags_raise():
    from rlcf_b200.params import build_parser
    from rlcf_b200.tune_cls_rl import _check_supported_flags
    base = ["DATA", "--tpt", "-a", "ViT-B/16"]
    _check_supported_flags(build_parser().parse_args(base))
    _check_supported_flags(build_parser().parse_args(base + ["--hard_aug", "1"]))            # host pre-augmentation
    for extra in (["--confidence_gap", "1"], ["--multiple_reward_models", "1"]):
        with pytest.raises(NotImplementedError):
            _check_supported_flags(build_parser().parse_args(base + extra))
    with pytest.raises(NotImplementedError):
        _check_supported_flags(build_parser().parse_args(["DATA", "-a", "ViT-B/16"]))        # no --tpt
    with pytest.raises(NotImplementedError):
        _check_supported_flags(build_parser().parse_args(["DATA", "--tpt"]))                 # default arch RN50


@pytest.mark.parametrize("name", ["tiny_prompt_middle", "tiny_prompt_front", "tiny_prompt_cls_word",
                                  "tiny_prompt_learned_cls", "tiny_prompt_rlcf_2step"])
def test_prompt_learner_layouts_match_the_reference(name):
    """PromptLearner's source map (class token at the end / in the middle / at the front, "[CLS]" inside ctx_init,
    learned class tokens; custom_clip.py:198-289) assembles exactly the prompt embeddings the reference's PromptLearner
    produced (tests/golden/*.npz `prompts0`, written by oracle/make_golden.py)."""
    import ast
    import numpy as np
    if _vocab_path() is None:
        pytest.skip("OpenAI BPE vocabulary not available on this machine")
    from rlcf_b200.clip import clip
    from rlcf_b200.clip.custom_clip import PromptLearner
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    cfg = ast.literal_eval(str(z["meta"]))
    if "prompts0" not in z.files:
        pytest.skip("fixture predates prompts0")
    model, _, _ = clip.load("synthetic:" + cfg["policy"] + ":0", device="cpu")
    names = ["tench", "goldfish", "great white shark", "tiger shark", "hammerhead", "electric ray", "stingray",
             "cock", "hen", "ostrich", "brambling", "goldfinch"][:cfg["C"]]
    learned = bool(cfg.get("learned_cls", False))
    pl = PromptLearner(model, names, None, n_ctx=4, ctx_init=cfg["ctx_init"], ctx_position=cfg.get("ctx_position", "end"),
                       learned_cls=learned, cls_init=torch.tensor(z["cls_init"]) if learned else None)
    assert np.array_equal(pl.tokenized_prompts.numpy(), z["tokens"])
    assert torch.equal(pl.ctx_init_state, torch.tensor(z["ctx_init"]))
    with torch.no_grad():
        assert torch.equal(pl(), torch.tensor(z["prompts0"]))
    src, ctx_pos, cls_pos = pl.source_map()
    for c in range(pl.n_cls):       # every context vector appears exactly once per class, where ctx_pos says
        for v in range(pl.n_ctx):
            assert src[c, ctx_pos[c, v]] == -1 - v and (src[c] == -1 - v).sum() == 1
    assert (cls_pos is not None) == learned
    flat = pl.learnable_flat()
    assert flat.numel() == (4 + (pl.n_cls if learned else 0)) * 128
    pl.load_flat(flat + 1.0)
    assert torch.equal(pl.learnable_flat(), flat + 1.0)
    pl.reset()
    assert torch.equal(pl.learnable_flat(), flat)


def test_bench_arms_build_the_same_config():
    """bench.py: the b200 arm and the reference arm describe the workload with the same `config` object -- identical
    except for images_per_step -- and the reference arm prefers the unmodified reference copy (kind "reference")."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for config in (2, 3, 5):
        wl = bench.workload_of(argparse.Namespace(config=config))
        a, b = bench.make_config(wl, "ln", config, 32, 1), bench.make_config(wl, "ln", config, 1, 1)
        assert set(a) == set(b)
        assert {k for k in a if a[k] != b[k]} == {"images_per_step"}
        assert a["workload"].startswith(wl["policy"]) and f"(config {config})" in a["workload"]
    assert bench.workload_of(argparse.Namespace(config=3))["tta_steps"] == 3
    from baseline import ref_harness
    if os.path.isdir("/root/reference/TPT"):
        from baseline import make_ref
        assert make_ref.make_ref() and ref_harness.available()
        import json
        with open(os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json")) as f:
            man = json.load(f)
        assert "tune_cls_rl.py" in man["files"] and "tpt_cls_rl.py" in man["files"] and "clip_reward.py" in man["files"]


def test_view_store_chunking(monkeypatch):
    """setup_view_store: the images of a step are split into equal chunks whose all-views activation store fits the
    budget; chunk_off turns view numbers inside a chunk into view numbers inside the step's batch."""
    import types
    import torch
    from rlcf_b200 import engine as E
    w = E.TowerWeights(kind="visual", d=64, heads=1, n_layers=3, L=5, E=8)
    w.ln_flat = torch.zeros(4)
    per_seq = E.ViewStore.bytes_per_seq(w)
    # x_pre + x_in[3] + x_mid[2] fp32, qkv (3d) + attn (d) fp16 for 2 layers, lse for 2 layers
    assert per_seq == 5 * 64 * (4 * 6 + 2 * 4 * 2) + 4 * 2 * 1 * 5
    B, V, S = 5, 4, 2
    run = types.SimpleNamespace(infer_row_stride=1)
    eng = types.SimpleNamespace()
    monkeypatch.setenv("RLCF_VIEW_STORE_GB", repr(2.5 * V * per_seq / 2 ** 30))      # room for two images
    E.setup_view_store(eng, w, run, B, V, S, torch.device("cpu"))
    assert eng.view_chunk == 2 and eng.views.n_seq == 2 * V and eng.views.n_layers == 2
    assert eng.chunk_off.tolist() == [0, 0, 0, 0, 8, 8, 8, 8, 16, 16]
    assert tuple(eng.views.x_in.shape) == (3, 2 * V * 5, 64) and eng.views.u is None
    monkeypatch.setenv("RLCF_VIEW_STORE_GB", "0")
    E.setup_view_store(eng, w, run, B, V, S, torch.device("cpu"))
    assert eng.views is None and eng.view_chunk == 0
    monkeypatch.setenv("RLCF_VIEW_STORE_GB", "48")
    E.setup_view_store(eng, w, types.SimpleNamespace(infer_row_stride=5), B, V, S, torch.device("cpu"))
    assert eng.views is None            # no class-token-only last block (RLCF_PRUNE_LAST=0): nothing to adopt from
