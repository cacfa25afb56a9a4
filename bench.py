#!/usr/bin/env python
"""bench.py -- adapted images/sec of the RLCF test-time-adaptation hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the UNMODIFIED reference (baseline/_ref) on the host CPU

Workload (BASELINE.json configs[1]): policy ViT-B/16, reward ViT-L/14, 64 views per image,
rho = 0.1 -> 6 selected views, K = 3 sampled classes, C = 200 classes, 1 TTA step, LayerNorm-only tuning,
synthetic 224x224 views and random-init CLIP weights (no datasets / checkpoints offline).
One "step" adapts `--images-per-step` (default 64) independent test images in one batched launch sequence (CUDA graph):
reset -> 64-view policy forward -> entropy selection -> reward forward on the selected views -> top-K/CLIPScore/
reward-weighted CE -> backward to the LayerNorm parameters -> AdamW -> adapted 1-view prediction.
Prints ONE JSON line on rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "adapted images/sec (ViT-B/16, 64 views, 1 TTA step)"
UNIT = "images/s"
WORKLOAD = dict(policy="ViT-B/16", reward="ViT-L/14", n_views=64, selection_p=0.1, sample_k=3, n_classes=200,
                tta_steps=1, lr=5e-3, mode="LayerNorm-only (--tune_norm 1)")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images-per-step", type=int, default=None,
                    help="independent test images adapted per launch sequence (default 64 for --mode ln, 8 otherwise)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 5],
                    help="BASELINE.json configs index: 2 = the metric's workload (default); 3 = same with 3 TTA steps; "
                         "5 = ViT-L/14 policy LN-tuning (informational)")
    ap.add_argument("--mode", default="ln", choices=["ln", "prompt", "full", "ret_i2t", "ret_t2i"],
                    help="ln = LayerNorm tuning (the BASELINE.json metric); prompt = prompt tuning; full = the whole "
                         "image encoder is trainable; ret_i2t / ret_t2i = retrieval TTA, BASELINE.json configs[3] "
                         "(all informational)")
    ap.add_argument("--queries-per-step", type=int, default=32, help="retrieval modes: queries adapted per launch sequence")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the bounded CPU-baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true",
                    help="skip the leg that times the UNMODIFIED reference (baseline/_ref: CLIPCLS_TTA + "
                         "test_time_tuning, fp16 autocast + GradScaler, nn.MultiheadAttention, one image at a time) on "
                         "the same GPU -- the denominator of north_star's 10x target")
    ap.add_argument("--no-other-modes", action="store_true",
                    help="skip the short informational runs of the other modes (prompt, full, config 3, config 5)")
    return ap.parse_args()


def workload_of(args) -> dict:
    wl = dict(WORKLOAD)
    if args.config == 3:
        wl["tta_steps"] = 3
    elif args.config == 5:
        wl["policy"] = "ViT-L/14"
    return wl


MODE_TEXT = {"ln": "LN-only", "prompt": "prompt tuning (ctx 4x512)", "full": "full image-encoder tuning"}
MODE_LONG = {"ln": WORKLOAD["mode"], "prompt": "prompt tuning (tpt_cls_rl.py, ctx_init a_photo_of_a)",
             "full": "full image-encoder tuning (--tune_norm 0, lr 1e-5)"}


def make_config(wl: dict, mode: str, config: int, images_per_step: int, world: int) -> dict:
    """The `config` object of the JSON line.  Both arms (--impl b200 / reference) build it here, so the two dicts are
    identical except for `images_per_step` (the CUDA path adapts a batch of independent images per launch sequence,
    the reference one image per call)."""
    return {"workload": "%s RLCF cls, %d views, %d step%s, reward %s (config %d), %s" % (
                wl["policy"], wl["n_views"], wl["tta_steps"], "" if wl["tta_steps"] == 1 else "s", wl["reward"], config,
                MODE_TEXT[mode]),
            **{k: v for k, v in wl.items() if k != "mode"}, "mode": MODE_LONG[mode],
            "images_per_step": images_per_step,
            "parallelism": f"dp{world} (independent images, no data-path collective)",
            "l2": "inputs larger than L2: every step reads a different batch of 64-view images (38.5 MB per image) "
                  "than the step before"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops_burst=p["bf16_tflops"], tflops_sustained=p["bf16_tflops_sustained"], hbm_gbs=p["hbm_gbs"],
                    source="MEASURED_PEAKS.json")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="B200_PROFILING.md fallback")


class NvmlSampler:
    """SM clock / power / throttle reasons sampled every 200 ms through NVML from a background thread (the same
    counters `nvidia-smi --query-gpu=clocks.sm,...,clocks_event_reasons.*` prints, without a second process)."""

    def __init__(self, gpu_index: int):
        import threading
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        uuid = torch.cuda.get_device_properties(gpu_index).uuid
        try:
            self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        except Exception:
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        self.stop_flag = threading.Event()
        self.sm, self.power, self.reasons = [], [], set()
        self.t = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        nv = self.nv
        masks = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, m in masks.items():
                    if r & m:
                        self.reasons.add(k)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def start(self):
        self.t.start()

    def stop(self) -> dict:
        self.stop_flag.set()
        self.t.join()
        mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "power_w_max": max(self.power) if self.power else None, "samples": len(sm),
                "reasons": sorted(self.reasons), "source": "nvml"}


class NullSampler:
    """RLCF_BENCH_SAMPLER=off: no sampling (only to measure what the sampling itself costs)."""

    def start(self):
        pass

    def stop(self):
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampling disabled"]}


def make_sampler(gpu_index: int):
    if os.environ.get("RLCF_BENCH_SAMPLER", "nvml") == "off":
        return NullSampler()
    if os.environ.get("RLCF_BENCH_SAMPLER", "nvml") == "nvml":
        try:
            return NvmlSampler(gpu_index)
        except Exception:
            pass
    return ClockSampler(gpu_index)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for n, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_harness():
    """baseline/ref_harness.py when the verbatim reference copy (baseline/_ref/TPT, made by baseline/make_ref.py in the
    build container; git-ignored, travels with the gpurun snapshot) is present, else None."""
    try:
        from baseline import ref_harness as H
    except Exception:
        return None
    return H if H.available() else None


def reference_cpu_images_per_s(wl, n_timed: int, n_warm: int):
    """(images/s, kind, cores, note): the reference's own per-image loop on all host cores, fp32.  kind "reference" =
    the unmodified reference code (tune_cls_rl.test_time_adapt_eval -> tpt_cls_rl.test_time_tuning -> CLIPCLS_TTA /
    CLIPRewards); kind "port" = the oracle restatement, used only when baseline/_ref is absent."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H = _ref_harness()
    from rlcf_b200 import synthetic as S
    views = S.make_views(2, wl["n_views"], 224, 11)
    if H is not None:
        run = H.ReferenceRun("cpu", wl, S.make_state_dict)
        try:
            if n_warm:
                run.run(views, n_warm)
            dt = run.run(views, n_timed)
        finally:
            run.close()
        return n_timed / dt, "reference", cores, "unmodified reference code from baseline/_ref/TPT"
    from oracle import rlcf_oracle as O
    sd_p, sd_r = O.make_clip_state_dict(wl["policy"], 0), O.make_clip_state_dict(wl["reward"], 1)
    cf = O.class_features(sd_p, O.make_tokens(wl["n_classes"], 49408))
    rc = O.class_features(sd_r, O.make_tokens(wl["n_classes"], 49408))
    cfg = O.OracleConfig(n_views=wl["n_views"], selection_p=wl["selection_p"], tta_steps=wl["tta_steps"],
                         sample_k=wl["sample_k"], lr=wl["lr"])
    V = wl["n_views"]
    for i in range(n_warm):
        O.adapt_one_image(sd_p, cf, views[(i % 2) * V:(i % 2 + 1) * V], cfg, sd_r, rc)
    t0 = time.perf_counter()
    for i in range(n_timed):
        O.adapt_one_image(sd_p, cf, views[(i % 2) * V:(i % 2 + 1) * V], cfg, sd_r, rc)
    dt = time.perf_counter() - t0
    return n_timed / dt, "port", cores, "oracle port (baseline/_ref absent)"


def run_reference(args):
    """The reference on the host CPU (fp32, all host threads), one image per step, through its own per-image loop."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    K = args.steps if args.steps is not None else 2
    W = args.warmup if args.warmup is not None else 1
    wl = workload_of(args)
    value, kind, cores, note = reference_cpu_images_per_s(wl, K, W)
    sample = (f"{K} images x full config-{args.config} sizes ({wl['n_views']} views, {wl['policy']} policy + "
              f"{wl['reward']} reward), one image per step, {W} warm-up image(s); {note}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": make_config(wl, args.mode if args.mode in MODE_TEXT else "ln", args.config, 1, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(wl):
    """Bounded CPU baseline on rank 0: two full-size images through the reference's own loop (10-15 s of host work)."""
    value, kind, cores, note = reference_cpu_images_per_s(wl, 2, 0)
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"2 images at full sizes ({wl['n_views']} views, {wl['policy']} policy + {wl['reward']} reward), "
                      f"no warm-up; {note}"}


def torch_gpu_baseline_leg(dev, wl, n_images=20, n_warm=3):
    """North_star's 10x denominator: the UNMODIFIED reference on the same GPU, as it runs there -- eager PyTorch,
    torch.cuda.amp.autocast fp16 + GradScaler(1000), nn.MultiheadAttention, one image per iteration, backward through
    all 64 views (TPT/tune_cls_rl.py:87,183-256, tpt_cls_rl.py:47-79).  None of this repo's kernels is involved."""
    H = _ref_harness()
    if H is None:
        return {"unavailable": "baseline/_ref/TPT missing (run python baseline/make_ref.py in the build container)"}
    from rlcf_b200 import synthetic as S
    run = H.ReferenceRun(str(dev), wl, S.make_state_dict)
    try:
        views = S.make_views(2, wl["n_views"], 224, 11)
        run.run(views, n_warm)
        dt = run.run(views, n_images)
    finally:
        run.close()
    return {"value": n_images / dt, "unit": UNIT, "kind": "reference",
            "what": "unmodified reference (baseline/_ref/TPT: tune_cls_rl.test_time_adapt_eval, CLIPCLS_TTA, CLIPRewards, "
                    "tpt_cls_rl.test_time_tuning) on cuda, eager PyTorch, fp16 autocast + GradScaler(1000)",
            "sample": f"{n_images} images, one per iteration (H2D of the 64 views included, as in the reference loop), "
                      f"after {n_warm} warm-up images"}


# ------------------------------------------------------------------------------------------------ CUDA arm
def build_engine(mode: str, wl: dict, B: int, dev, reward_seed: int = 1):
    """Policy + reward towers with synthetic weights and the batched engine of the requested mode."""
    from rlcf_b200 import engine as E, synthetic as S
    sd_p = S.make_state_dict(wl["policy"], 0, dev)
    sd_r = S.make_state_dict(wl["reward"], reward_seed, dev)
    rew = E.prepare_visual(sd_r)
    tok = S.make_tokens(wl["n_classes"], 49408)
    rc = E.text_features(E.prepare_text(sd_r), tok)
    logit_scale = float(sd_p["logit_scale"].exp())
    cfg = E.RlcfConfig(n_views=wl["n_views"], selection_p=wl["selection_p"], tta_steps=wl["tta_steps"],
                       sample_k=wl["sample_k"], lr=wl["lr"])
    if mode == "prompt":
        tok[:, 1:5] = torch.tensor([320, 1125, 539, 320])   # "a photo of a": the 4 context positions
        ctx_init = sd_p["token_embedding.weight"][tok[0, 1:5].to(dev)]
        return E.PromptEngine(E.prepare_visual(sd_p), E.prepare_text(sd_p, need_grad=True), tok, ctx_init, logit_scale,
                              cfg, B, reward=rew, reward_class_feat=rc)
    if mode == "full":
        from rlcf_b200 import full_tune as FT
        cf = E.text_features(E.prepare_text(sd_p), tok)
        cfg.lr = 1e-5                                        # scripts/rlcf-tune.sh
        return FT.FullTuneEngine(sd_p, cf, logit_scale, cfg, B, rew, rc)
    pol = E.prepare_visual(sd_p, need_grad=True)
    cf = E.text_features(E.prepare_text(sd_p), tok)
    return E.RlcfEngine(pol, cf, logit_scale, cfg, B, reward=rew, reward_class_feat=rc)


def measure(eng, wl, B, K, W, dev, rank, world, dist, use_graph=True, warm_seconds=2.0):
    """Times K steps of the device-resident path (`value`) and K steps of the host-buffer path (`e2e`).  A step =
    adapt B images (reset -> ... -> adapted prediction) + the top-1/top-5 counters of tools.accuracy, all of it this
    library's kernels (no eager torch op inside either timed region)."""
    from rlcf_b200 import _lib, ops, synthetic as S
    V, C = wl["n_views"], wl["n_classes"]
    # two different resident input batches, alternated: 2 x B x 38.5 MB (> 126 MB L2 for B >= 2)
    batches = [S.make_views(B, V, 224, 1000 + 17 * rank + i, device=dev) for i in range(2)]
    labels = [torch.randint(0, C, (B,), device=dev, dtype=torch.int64) for _ in range(2)]
    hits = torch.zeros(3, device=dev, dtype=torch.int64)
    in_bytes = batches[0].numel() * 4

    l0 = _lib.launch_count()
    if use_graph:
        eng.capture(batches[0])
        launches_per_step = (_lib.launch_count() - l0) // 3 + 1   # 2 eager warm-ups + 1 capture; + accuracy_count
        adapt = eng.adapt_graph
    else:
        eng.adapt(batches[0])
        launches_per_step = _lib.launch_count() - l0 + 1
        adapt = eng.adapt

    def step(i):
        ops.accuracy_count(adapt(batches[i % 2]), labels[i % 2], hits)

    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    # extra untimed warm-up until ~2 s of work have run, so that clocks / power state are in their steady regime
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < warm_seconds:
        step(0)
        step(1)
        torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    local = dev.index or 0
    sampler = make_sampler(local)
    hits.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    counters = hits.clone()
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)   # the only collective: final accuracy counters
    ms_total = float(ms.item())
    value = world * B * K / (ms_total / 1e3)

    # ---------------- end-to-end with host buffers (`e2e`): pinned H2D of every step's views and labels, D2H of the
    # adapted logits and of the counters
    host_in = [b.cpu().pin_memory() for b in batches]
    host_lab = [l.cpu().pin_memory() for l in labels]
    lab_dev = torch.empty_like(labels[0])
    host_out = torch.empty(B, C, dtype=torch.float32).pin_memory()
    host_hits = torch.zeros(3, dtype=torch.int64).pin_memory()
    pipe = eng.host_pipeline() if (use_graph and hasattr(eng, "host_pipeline")) else None

    def e2e_step(i):
        if pipe is not None:
            if i + 1 < K:
                pipe.submit(host_in[(i + 1) % 2], (i + 1) % 2)
            lab_dev.copy_(host_lab[i % 2], non_blocking=True)
            pipe.run(i % 2, host_out)
        else:
            lab_dev.copy_(host_lab[i % 2], non_blocking=True)
            host_out.copy_(eng.adapt(host_in[i % 2].to(dev, non_blocking=True)), non_blocking=True)
        ops.accuracy_count(eng.logits_final, lab_dev, hits)
        host_hits.copy_(hits, non_blocking=True)

    if pipe is not None:
        # warm the copy path and bring clocks / power back to their steady regime (pinning the host buffers above
        # left the GPU idle for seconds; timing right after would measure a boost transient, not throughput)
        t_warm = time.perf_counter()
        while time.perf_counter() - t_warm < warm_seconds:
            pipe.submit(host_in[0], 0)
            pipe.submit(host_in[1], 1)
            pipe.run(0, host_out)
            pipe.run(1, host_out)
            torch.cuda.synchronize()
    barrier()
    e0.record()
    if pipe is not None:
        # software pipeline: the H2D copy of step i+1 (side stream) overlaps the adaptation of step i;
        # every step's views are copied from pinned host memory inside this timed region
        pipe.submit(host_in[0], 0)
    for i in range(K):
        e2e_step(i)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (float(ms2.item()) / 1e3)
    return dict(value=value, ms_total=ms_total, e2e_value=e2e_value, clocks=clocks, counters=counters,
                launches_per_step=launches_per_step, in_bytes=in_bytes + B * 8, out_bytes=host_out.numel() * 4 + 24,
                batches=batches)


def gemm_roofline(eng, batch, ms_per_step, B, value, world, args):
    """Roofline of the dominant kernel (the tcgen05 GEMM), every launch of one step timed live with CUDA events."""
    from rlcf_b200 import ops
    pk = peaks()
    ops.GEMM_TIMER = []
    eng.adapt(batch)
    torch.cuda.synchronize()
    recs, ops.GEMM_TIMER = ops.GEMM_TIMER, None
    g_ms = sum(a.elapsed_time(b) for (_, _, _, a, b) in recs)
    g_flops = sum(2.0 * m * n * k for (m, n, k, _, _) in recs)
    achieved = g_flops / (g_ms / 1e3) / 1e12
    flops_img = eng.algorithmic_flops_per_image()
    step_tflops = flops_img * value / world / 1e12
    traffic, traffic_src = None, None
    for name in ("r2_gemm_traffic_b%d.json" % B, {16: "r1_gemm_traffic.json", 32: "r1_gemm_traffic_b32.json"}.get(B, "none")):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath) and args.mode == "ln" and args.config == 2:   # DRAM bytes per GEMM launch, committed ncu capture of this workload
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj["dram_bytes_per_launch_avg"], tj["source"]
            break
    return {
        "bound": "tensor", "kernel": "gemm_f16_kernel (tcgen05/TMA, all %d launches of one step)" % len(recs),
        "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"],
        "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": pk["source"] + " (sustained bf16 dense, of measured)",
        "avg_launch_us": 1e3 * g_ms / len(recs), "gemm_share_of_step": g_ms / ms_per_step,
        "flops_per_launch": g_flops / len(recs),
        # algorithmic = what the algorithm needs: the last block of an inference forward only feeds the class token into
        # ln_post, so its out_proj / MLP / attention rows of the other tokens are not computed (and not counted);
        # reference_gflop_per_image = SURVEY.md 8(d)'s count with every block on every token, as the reference runs it
        "whole_step": {"algorithmic_gflop_per_image": flops_img / 1e9,
                       "reference_gflop_per_image": (eng.reference_flops_per_image() / 1e9
                                                     if hasattr(eng, "reference_flops_per_image") else None),
                       "achieved_tflops": step_tflops, "frac": step_tflops / pk["tflops_sustained"]},
    }


def other_modes_leg(dev, rank):
    """Short informational runs of the other supported modes at the same shapes, so that the driver (not only the
    builder) records them with clocks: prompt tuning, full image-encoder tuning, config 3 (3 TTA steps), config 5
    (ViT-L/14 policy).  3 timed steps each after 3 warm-up steps + 1 s of steady-state warm-up."""
    import gc
    pk = peaks()
    out = {}
    for name, mode, config in (("prompt", "prompt", 2), ("full", "full", 2), ("config3_ln", "ln", 3),
                               ("config5_ln", "ln", 5)):
        try:
            wl = workload_of(argparse.Namespace(config=config))
            B = 16      # short runs: a quarter of the headline's 64 images per step (config 3 at 64: profiles/r2_bench_config3_b64.json)
            eng = build_engine(mode, wl, B, dev, reward_seed=3 if config == 5 else 1)
            m = measure(eng, wl, B, 3, 3, dev, rank, 1, None, warm_seconds=1.0)
            fl = eng.algorithmic_flops_per_image()
            out[name] = {"value": m["value"], "e2e": m["e2e_value"], "unit": UNIT, "images_per_step": B, "steps": 3,
                         "ms_per_step": m["ms_total"] / 3, "algorithmic_gflop_per_image": fl / 1e9,
                         "frac_of_sustained_tensor_peak": fl * m["value"] / 1e12 / pk["tflops_sustained"],
                         "workload": make_config(wl, mode, config, B, 1)["workload"],
                         "clocks": {k: m["clocks"].get(k) for k in ("sm_mhz", "reasons")}}
            if hasattr(eng, "text_tokens"):     # prompt tuning: positions of CLIP's 77 the causal text tower runs on
                out[name]["text_positions"] = f"{eng.text_tokens} of 77 (up to the last EOT; the padding cannot reach it)"
            del eng, m
        except Exception as e:   # informational leg: never take the headline line down with it
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        gc.collect()
        torch.cuda.empty_cache()
    return out


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rlcf_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from rlcf_b200 import _lib

    K = args.steps if args.steps is not None else 10
    W = max(3, args.warmup if args.warmup is not None else 3)
    # 64 images per launch sequence: measured 202 / 209 / 213 / 212 images/s at 16 / 32 / 48 / 64 (one box, same run)
    B = args.images_per_step if args.images_per_step is not None else (64 if args.mode == "ln" else 8)
    wl = workload_of(args)
    eng = build_engine(args.mode, wl, B, dev, reward_seed=1 if args.config != 5 else 3)
    m = measure(eng, wl, B, K, W, dev, rank, world, dist, use_graph=not args.no_graph)
    value, ms_total = m["value"], m["ms_total"]
    roofline = gemm_roofline(eng, m["batches"][0], ms_total / K, B, value, world, args)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    hits = m["counters"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": make_config(wl, args.mode, args.config, B, world),
        "impl_details": {"cuda_graph": not args.no_graph, "gemm_cta_group": _lib.set_gemm_cta_group(0),
                         "resident_input_bytes": 2 * (m["in_bytes"] - B * 8),
                         "hbm_peak_allocated_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1),
                         "view_store": (None if getattr(eng, "views", None) is None else
                                        {"images_per_chunk": eng.view_chunk, "gb": round(eng.view_store_gb, 1)})},
        "clocks": m["clocks"],
        "e2e": {"value": m["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": m["in_bytes"],
                "d2h_bytes_per_step": m["out_bytes"]},
        "gpu_launches": int(m["launches_per_step"] * K),
        "roofline": roofline,
        "accuracy_counters": {"top1_hits": int(hits[0]), "top5_hits": int(hits[1]), "count": int(hits[2]),
                              "note": "random labels on synthetic data; rlcf_accuracy_count on the device, summed "
                                      "over ranks with one NCCL all-reduce"},
    }
    del eng, m
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if world == 1 and not args.no_torch_gpu_baseline and args.mode == "ln":
        try:
            line["torch_gpu_baseline"] = torch_gpu_baseline_leg(dev, wl)
            if "value" in line["torch_gpu_baseline"]:
                line["torch_gpu_baseline"]["e2e_speedup_vs_it"] = line["e2e"]["value"] / line["torch_gpu_baseline"]["value"]
        except Exception as e:
            line["torch_gpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        gc.collect()
        torch.cuda.empty_cache()
    if world == 1 and not args.no_other_modes and args.mode == "ln" and args.config == 2:
        line["other_modes"] = other_modes_leg(dev, rank)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(wl)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ retrieval (config 4)
def run_retrieval(args):
    """BASELINE.json configs[3]: retrieval/clip_ret_policy.py, ViT-B/16 policy + ViT-L/14 reward, COCO shape
    (5 000 images x 25 000 captions), 8 TTA steps per query, every parameter of the query's encoder tuned.
    Informational line (the headline metric is the classification loop): queries/s and the achieved fraction of
    the HBM roofline -- with M = 197 (or 77) rows per weight group the path is weight-bandwidth bound."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rlcf_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from rlcf_b200 import _lib, engine as E, retrieval as R, synthetic as S
    i2t = args.mode == "ret_i2t"
    K = args.steps if args.steps is not None else 5
    W = max(3, args.warmup if args.warmup is not None else 3)
    Q = args.queries_per_step
    n_gallery = 25000 if i2t else 5000
    sd_p = S.make_state_dict("ViT-B/16", 0, dev)
    sd_r = S.make_state_dict("ViT-L/14", 1, dev)
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    gal_p = torch.nn.functional.normalize(torch.randn(n_gallery, 512, generator=g, device=dev), dim=-1)
    gal_r = torch.nn.functional.normalize(torch.randn(n_gallery, 768, generator=g, device=dev), dim=-1)
    rcfg = R.RetrievalConfig(tta_steps=8, sample_k=20 if i2t else 12, lr=1e-6)
    if i2t:
        eng = R.ImageQueryEngine(sd_p, gal_p, float(sd_p["logit_scale"].exp()), rcfg, Q, E.prepare_visual(sd_r), gal_r)
        batches = [S.make_views(Q, 1, 224, 2000 + 17 * rank + i, device=dev) for i in range(2)]
    else:
        eng = R.TextQueryEngine(sd_p, gal_p, rcfg, Q, E.prepare_text(sd_r), gal_r)
        batches = [S.make_tokens(Q, 49408, seed=31 + 17 * rank + i).to(dev) for i in range(2)]
    del sd_p, sd_r
    l0 = _lib.launch_count()
    eng.adapt(batches[0])
    launches_per_step = _lib.launch_count() - l0
    step = eng.adapt
    if not args.no_graph:
        eng.capture(batches[0])
        step = eng.adapt_graph
    for i in range(W):
        step(batches[i % 2])
    torch.cuda.synchronize()
    sampler = make_sampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0.record()
    for i in range(K):
        step(batches[i % 2])
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * Q * K / (ms_total / 1e3)
    # end to end: queries from pinned host memory, score rows back to pinned host memory
    host_in = [b.cpu().pin_memory() for b in batches]
    host_out = torch.empty(Q, n_gallery, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        host_out.copy_(step(host_in[i % 2].to(dev, non_blocking=True)), non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * Q * K / (float(ms2.item()) / 1e3)
    pk = peaks()
    gbs = eng.bytes_per_query() * (value / world) / 1e9
    if rank == 0:
        line = {
            "metric": "adapted retrieval queries/sec (ViT-B/16 policy, ViT-L/14 reward, 8 TTA steps, all encoder "
                      "parameters tuned)", "value": value, "unit": "queries/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": "retrieval %s, COCO shape (config 4): %d gallery candidates, K=%d, 8 steps, lr 1e-6"
                                   % ("image->text" if i2t else "text->image", n_gallery, rcfg.sample_k),
                       "queries_per_step": Q, "parallelism": f"dp{world} (independent queries)",
                       "cuda_graph": bool(not args.no_graph),
                       "l2": "per-query weights + Adam state: %.1f GB per step, far beyond L2" % (
                           Q * eng.lay.total * 16 / 1e9)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": host_in[0].numel() * host_in[0].element_size(),
                    "d2h_bytes_per_step": host_out.numel() * 4},
            "gpu_launches": int(launches_per_step * K),
            "roofline": {"bound": "hbm", "kernel": "whole step (per-query weight streaming: grouped GEMMs, wgrad, AdamW, casts)",
                         "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "traffic": None, "algorithmic_bytes_per_query": eng.bytes_per_query(),
                         "peak_source": pk["source"]},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode in ("ret_i2t", "ret_t2i"):
        run_retrieval(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
