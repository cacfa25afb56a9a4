#!/usr/bin/env python
"""bench.py -- adapted images/sec of the RLCF test-time-adaptation hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)

Workload (BASELINE.json configs[1]): policy ViT-B/16, reward ViT-L/14, 64 views per image,
rho = 0.1 -> 6 selected views, K = 3 sampled classes, C = 200 classes, 1 TTA step, LayerNorm-only tuning,
synthetic 224x224 views and random-init CLIP weights (no datasets / checkpoints offline).
One "step" adapts `--images-per-step` (default 32) independent test images in one batched launch sequence (CUDA graph):
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
                    help="independent test images adapted per launch sequence (default 32 for --mode ln, 8 otherwise)")
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
    ap.add_argument("--torch-gpu-baseline", action="store_true",
                    help="also time the reference algorithm in plain PyTorch on the GPU (fp16 autocast + GradScaler, "
                         "one image at a time, 64-view backward) -- the denominator of north_star's 10x target")
    return ap.parse_args()


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
def run_reference(args):
    """The reference algorithm (CPU, fp32, all host threads) through the oracle port: one image per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import rlcf_oracle as O
    K = args.steps if args.steps is not None else 2
    W = args.warmup if args.warmup is not None else 1
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOAD
    sd_p = O.make_clip_state_dict(wl["policy"], 0)
    sd_r = O.make_clip_state_dict(wl["reward"], 1)
    cf = O.class_features(sd_p, O.make_tokens(wl["n_classes"], 49408))
    rc = O.class_features(sd_r, O.make_tokens(wl["n_classes"], 49408))
    cfg = O.OracleConfig(n_views=wl["n_views"], selection_p=wl["selection_p"], tta_steps=wl["tta_steps"],
                         sample_k=wl["sample_k"], lr=wl["lr"])
    views = O.make_views(1, wl["n_views"], 224, 11)
    for _ in range(W):
        O.adapt_one_image(sd_p, cf, views, cfg, sd_r, rc)
    t0 = time.perf_counter()
    for _ in range(K):
        O.adapt_one_image(sd_p, cf, views, cfg, sd_r, rc)
    dt = time.perf_counter() - t0
    value = K / dt
    sample = f"{K} images x full config-2 sizes (64 views, B/16 policy + L/14 reward), one image per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ViT-B/16 RLCF cls, 64 views, 1 step, reward ViT-L/14 (config 2), LN-only", **wl,
                   "images_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg():
    """Bounded CPU baseline on rank 0: the oracle port on two full-size images (about 15 s of host work on 16 cores)."""
    from oracle import rlcf_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOAD
    sd_p = O.make_clip_state_dict(wl["policy"], 0)
    sd_r = O.make_clip_state_dict(wl["reward"], 1)
    cf = O.class_features(sd_p, O.make_tokens(wl["n_classes"], 49408))
    rc = O.class_features(sd_r, O.make_tokens(wl["n_classes"], 49408))
    cfg = O.OracleConfig(n_views=wl["n_views"], selection_p=wl["selection_p"], tta_steps=wl["tta_steps"],
                         sample_k=wl["sample_k"], lr=wl["lr"])
    n_img, V = 2, wl["n_views"]
    views = O.make_views(n_img, V, 224, 11)
    t0 = time.perf_counter()
    for i in range(n_img):
        O.adapt_one_image(sd_p, cf, views[i * V:(i + 1) * V], cfg, sd_r, rc)
    dt = time.perf_counter() - t0
    return {"value": n_img / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_img} images at full config-2 sizes (64 views, 6 selected, B/16 policy + L/14 reward), no warm-up"}


def torch_gpu_baseline_leg(dev, n_images=10):
    """The reference algorithm as the reference runs it on a GPU: eager PyTorch, torch.cuda.amp.autocast fp16 +
    GradScaler(1000), one image per iteration, backward through all 64 views (TPT/tune_cls_rl.py:87,192-222) --
    via the oracle port moved to the device.  Reported baseline only; none of this repo's kernels are involved."""
    from oracle import rlcf_oracle as O
    wl = WORKLOAD
    sd_p = O.make_clip_state_dict(wl["policy"], 0)
    sd_r = O.make_clip_state_dict(wl["reward"], 1)
    cf = O.class_features(sd_p, O.make_tokens(wl["n_classes"], 49408)).to(dev)
    rc = O.class_features(sd_r, O.make_tokens(wl["n_classes"], 49408)).to(dev)
    sd_p = {k: v.to(dev) for k, v in sd_p.items() if k.startswith("visual.") or k == "logit_scale"}
    sd_r = {k: v.to(dev) for k, v in sd_r.items() if k.startswith("visual.") or k == "logit_scale"}
    cfg = O.OracleConfig(n_views=wl["n_views"], selection_p=wl["selection_p"], tta_steps=wl["tta_steps"],
                         sample_k=wl["sample_k"], lr=wl["lr"])
    views = O.make_views(2, wl["n_views"], 224, 11).to(dev)
    scaler = torch.amp.GradScaler("cuda", init_scale=1000)
    V = wl["n_views"]
    for i in range(3):
        O.adapt_one_image(sd_p, cf, views[(i % 2) * V:(i % 2 + 1) * V], cfg, sd_r, rc, amp=True, scaler=scaler)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_images):
        O.adapt_one_image(sd_p, cf, views[(i % 2) * V:(i % 2 + 1) * V], cfg, sd_r, rc, amp=True, scaler=scaler)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": n_images / dt, "unit": UNIT, "kind": "oracle port on cuda, eager PyTorch, fp16 autocast + GradScaler",
            "sample": f"{n_images} images, one per iteration, after 3 warm-up images"}


# ------------------------------------------------------------------------------------------------ CUDA arm
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
    from rlcf_b200 import _lib, engine as E, ops, synthetic as S

    K = args.steps if args.steps is not None else 10
    W = max(3, args.warmup if args.warmup is not None else 3)
    B = args.images_per_step if args.images_per_step is not None else (32 if args.mode == "ln" else 8)
    wl = dict(WORKLOAD)
    if args.config == 3:
        wl["tta_steps"] = 3
    elif args.config == 5:
        wl["policy"] = "ViT-L/14"
    sd_p = S.make_state_dict(wl["policy"], 0, dev)
    sd_r = S.make_state_dict(wl["reward"], 1 if args.config != 5 else 3, dev)
    rew = E.prepare_visual(sd_r)
    tok = S.make_tokens(wl["n_classes"], 49408)
    rc = E.text_features(E.prepare_text(sd_r), tok)
    logit_scale = float(sd_p["logit_scale"].exp())
    cfg = E.RlcfConfig(n_views=wl["n_views"], selection_p=wl["selection_p"], tta_steps=wl["tta_steps"],
                       sample_k=wl["sample_k"], lr=wl["lr"])
    if args.mode == "prompt":
        tok[:, 1:5] = torch.tensor([320, 1125, 539, 320])   # "a photo of a": the 4 context positions
        ctx_init = sd_p["token_embedding.weight"][tok[0, 1:5].to(dev)]
        eng = E.PromptEngine(E.prepare_visual(sd_p), E.prepare_text(sd_p, need_grad=True), tok, ctx_init, logit_scale,
                             cfg, B, reward=rew, reward_class_feat=rc)
    elif args.mode == "full":
        from rlcf_b200 import full_tune as FT
        cf = E.text_features(E.prepare_text(sd_p), tok)
        cfg.lr = 1e-5                                        # scripts/rlcf-tune.sh
        eng = FT.FullTuneEngine(sd_p, cf, logit_scale, cfg, B, rew, rc)
    else:
        pol = E.prepare_visual(sd_p, need_grad=True)
        cf = E.text_features(E.prepare_text(sd_p), tok)
        eng = E.RlcfEngine(pol, cf, logit_scale, cfg, B, reward=rew, reward_class_feat=rc)
    del sd_p, sd_r
    V = wl["n_views"]
    # two different resident input batches, alternated: 2 x B x 38.5 MB (> 126 MB L2 for B >= 2)
    batches = [S.make_views(B, V, 224, 1000 + 17 * rank + i, device=dev) for i in range(2)]
    labels = torch.randint(0, wl["n_classes"], (B,), device=dev)
    in_bytes = batches[0].numel() * 4

    l0 = _lib.launch_count()
    if args.no_graph:
        step = eng.adapt
        eng.adapt(batches[0])
        launches_per_step = _lib.launch_count() - l0
    else:
        eng.capture(batches[0])
        launches_per_step = (_lib.launch_count() - l0) // 3   # 2 eager warm-ups + 1 capture
        step = eng.adapt_graph
    for i in range(W):
        step(batches[i % 2])
    torch.cuda.synchronize()
    # extra untimed warm-up until ~2 s of work have run, so that clocks / power state are in their steady regime
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 2.0:
        step(batches[0])
        step(batches[1])
        torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    sampler = make_sampler(local)
    hits = torch.zeros(3, device=dev, dtype=torch.int64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record()
    for i in range(K):
        logits = step(batches[i % 2])
        # accuracy counters as tools.accuracy (TPT/utils/tools.py:84-98): top-1 / top-5 hits, count
        top5 = logits.topk(5, dim=1).indices
        hits[0] += (top5[:, 0] == labels).sum()
        hits[1] += (top5 == labels[:, None]).any(1).sum()
        hits[2] += B
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(hits, op=dist.ReduceOp.SUM)   # the only collective: final accuracy counters
    ms_total = float(ms.item())
    value = world * B * K / (ms_total / 1e3)

    # ---------------- end-to-end with host buffers (`e2e`): pinned H2D of every step's views + D2H of the logits
    host_in = [b.cpu().pin_memory() for b in batches]
    host_out = torch.empty(B, wl["n_classes"], dtype=torch.float32).pin_memory()
    pipe = None if (args.no_graph or not hasattr(eng, "host_pipeline")) else eng.host_pipeline()
    if pipe is not None:
        # warm the copy path and bring clocks / power back to their steady regime (pinning the host buffers above
        # left the GPU idle for seconds; timing right after would measure a boost transient, not throughput)
        t_warm = time.perf_counter()
        while time.perf_counter() - t_warm < 2.0:
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
            if i + 1 < K:
                pipe.submit(host_in[(i + 1) % 2], (i + 1) % 2)
            pipe.run(i % 2, host_out)
    else:
        for i in range(K):
            host_out.copy_(eng.adapt(host_in[i % 2].to(dev, non_blocking=True)), non_blocking=True)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (float(ms2.item()) / 1e3)

    # ---------------- roofline of the dominant kernel (the tcgen05 GEMM), timed live per launch with CUDA events
    pk = peaks()
    ops.GEMM_TIMER = []
    eng.adapt(batches[0])
    torch.cuda.synchronize()
    recs, ops.GEMM_TIMER = ops.GEMM_TIMER, None
    g_ms = sum(a.elapsed_time(b) for (_, _, _, a, b) in recs)
    g_flops = sum(2.0 * m * n * k for (m, n, k, _, _) in recs)
    achieved = g_flops / (g_ms / 1e3) / 1e12
    flops_img = eng.algorithmic_flops_per_image()
    step_tflops = flops_img * value / world / 1e12
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", {16: "r1_gemm_traffic.json", 32: "r1_gemm_traffic_b32.json"}.get(B, "none"))
    if os.path.exists(tpath) and args.mode == "ln" and args.config == 2:   # DRAM bytes per GEMM launch from the committed ncu capture of this workload
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["dram_bytes_per_launch_avg"], tj["source"]
    roofline = {
        "bound": "tensor", "kernel": "gemm_f16_kernel (tcgen05/TMA, all %d launches of one step)" % len(recs),
        "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"],
        "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": pk["source"] + " (sustained bf16 dense, of measured)",
        "avg_launch_us": 1e3 * g_ms / len(recs), "gemm_share_of_step": g_ms / (ms_total / K),
        "flops_per_launch": g_flops / len(recs),
        "whole_step": {"algorithmic_gflop_per_image": flops_img / 1e9, "achieved_tflops": step_tflops,
                       "frac": step_tflops / pk["tflops_sustained"]},
    }
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": "%s RLCF cls, 64 views, %d step(s), reward ViT-L/14 (config %d), " % (
                        wl["policy"], wl["tta_steps"], args.config)
                               + {"ln": "LN-only", "prompt": "prompt tuning (ctx 4x512)",
                                  "full": "full image-encoder tuning"}[args.mode], **wl,
                   "mode": {"ln": wl["mode"], "prompt": "prompt tuning (tpt_cls_rl.py, ctx_init a_photo_of_a)",
                            "full": "full image-encoder tuning (--tune_norm 0, lr 1e-5)"}[args.mode],
                   "images_per_step": B, "parallelism": f"dp{world} (independent images, no data-path collective)",
                   "l2": "inputs larger than L2: two alternating resident batches of %.0f MB" % (in_bytes / 1e6),
                   "cuda_graph": not args.no_graph, "gemm_cta_group": _lib.set_gemm_cta_group(0)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": host_out.numel() * 4},
        "gpu_launches": int(launches_per_step * K),
        "roofline": roofline,
        "accuracy_counters": {"top1_hits": int(hits[0]), "top5_hits": int(hits[1]), "count": int(hits[2]),
                              "note": "random labels on synthetic data; summed over ranks with one NCCL all-reduce"},
    }
    if world == 1 and args.torch_gpu_baseline:
        del eng
        torch.cuda.empty_cache()
        line["torch_gpu_baseline"] = torch_gpu_baseline_leg(dev)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg()
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
