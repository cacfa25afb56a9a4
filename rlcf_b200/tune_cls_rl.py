"""RLCF evaluation driver for image-encoder (LayerNorm) tuning -- the loop of TPT/tune_cls_rl.py:183-256, batched.

The reference adapts one test image per iteration.  Every image restarts from the same weights and an empty Adam
state (tune_cls_rl.py:210-213), so images are independent: this driver gathers `images_per_step` samples from the
loader, adapts them in ONE batched launch sequence (CUDA graph) and scores each adapted prediction.  With
--momentum_update 1 the reset state depends on sample order (custom_clip.py:460-475) and the driver falls back to
the reference's one-at-a-time order.

Multi-GPU: one process per GPU (torchrun); rank r evaluates samples r, r+R, ...; the only collective is one
all-reduce of (top-1 hits, top-5 hits, count) per dataset.

    python -m rlcf_b200.tune_cls_rl --synthetic -a ViT-B/16 --reward_arch ViT-L/14 --tpt --tune_norm 1 \
        --batch_size 64 --selection_p 0.1 --tta_steps 1 --sample_k 3 --n_images 64
"""
from __future__ import annotations

import json
import os
import time
from copy import deepcopy

import torch

from . import datautils, synthetic
from .clip.custom_clip import CLIPCLS_TTA
from .clip_reward import get_reward_model
from .params import get_args
from .tpt_cls_rl import engine_config, test_time_tuning
from .utils.tools import AverageMeter, ProgressMeter, Summary, accuracy, set_random_seed


def _stack_views(images, device):
    """One loader sample -> [V,3,H,W] on the device (tune_cls_rl.py:194-207: list of [1,3,H,W] views or a tensor), or
    (uint8 image, ViewPlan) from ImageFolderViews: the views are then generated on the device."""
    if isinstance(images, tuple) and len(images) == 2 and isinstance(images[1], datautils.ViewPlan):
        return datautils.run_plan(images[0], images[1], device)
    if isinstance(images, (list, tuple)):
        return torch.cat([im.to(device, non_blocking=True) for im in images], dim=0)
    if images.dim() > 4:
        images = images.squeeze(0)
    return images.to(device, non_blocking=True)


def test_time_adapt_eval(val_loader, model, optimizer, optim_state, scaler, args, device=None, reward_model=None):
    """Returns [top1, top5] (percent, 3 decimals) over this rank's share of the loader, summed over ranks."""
    device = device if device is not None else torch.device("cuda", args.gpu)
    batch_time = AverageMeter("Time", ":6.3f", Summary.NONE)
    top1 = AverageMeter("Acc@1", ":6.2f", Summary.AVERAGE)
    top5 = AverageMeter("Acc@5", ":6.2f", Summary.AVERAGE)
    progress = ProgressMeter(len(val_loader), [batch_time, top1, top5], prefix="Test: ")
    B = 1 if getattr(args, "momentum_update", 0) else max(1, getattr(args, "images_per_step", 8))
    hits = torch.zeros(3, dtype=torch.int64, device=device)
    pend_views, pend_targets = [], []
    end = time.time()

    def flush():
        nonlocal end
        n = len(pend_views)
        if n == 0:
            return
        optimizer.load_state_dict(optim_state)                   # tune_cls_rl.py:213
        views = torch.cat(pend_views, dim=0)
        target = torch.cat(pend_targets, dim=0)
        model.train()
        if n == 1:
            model.reset()                                        # tune_cls_rl.py:210
            test_time_tuning(model, views, optimizer, scaler, args, reward_model=reward_model)
            model.eval()
            output = model(views[:1])                            # tune_cls_rl.py:220-222
            model.momentum_update_model()                        # tune_cls_rl.py:240
        else:
            cfg = engine_config(args, optimizer, reward_model)
            cfg.n_views = views.shape[0] // n
            model.reset()
            eng = model.engine(cfg, n, reward_model)
            if hasattr(eng, "init_params"):      # LayerNorm tuning: adapt from the model's current LN parameters
                eng.init_params.copy_(model.clip_model.visual.ln_flat())
            output = eng.adapt_graph(views.float().contiguous())
            model.eval()
        acc1, acc5 = accuracy(output, target, topk=(1, 5))       # tune_cls_rl.py:243-245
        hits[0] += torch.round(acc1[0] * n / 100).long()
        hits[1] += torch.round(acc5[0] * n / 100).long()
        hits[2] += n
        top1.update(acc1[0], n)
        top5.update(acc5[0], n)
        batch_time.update(time.time() - end)
        end = time.time()
        pend_views.clear()
        pend_targets.clear()

    n_cls = getattr(model, "n_cls", None)
    for i, (images, target) in enumerate(val_loader):
        if n_cls is not None and int(target.view(-1)[0]) >= n_cls:      # host-side label, no device sync
            raise RuntimeError(f"label {int(target.view(-1)[0])} outside the model's {n_cls} classes: call "
                               "model.reset_classnames_and_state(...) for this test set (tune_cls_rl.py:139)")
        pend_views.append(_stack_views(images, device))
        pend_targets.append(target.to(device, non_blocking=True).view(-1)[:1])
        if len(pend_views) == B:
            flush()
        if (i + 1) % args.print_freq == 0:
            progress.display(i)
    while pend_views:       # ragged tail: adapt the leftovers one image at a time
        tail_v, tail_t = pend_views[1:], pend_targets[1:]
        del pend_views[1:], pend_targets[1:]
        flush()
        pend_views.extend(tail_v)
        pend_targets.extend(tail_t)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(hits, op=torch.distributed.ReduceOp.SUM)
    progress.display_summary()
    n = max(1, int(hits[2]))
    return [round(100.0 * int(hits[0]) / n, 3), round(100.0 * int(hits[1]) / n, 3)]


class SyntheticViews(torch.utils.data.Dataset):
    """Seeded synthetic samples: ([V,3,res,res] views, label); rank r of R sees indices r, r+R, ..."""

    def __init__(self, n_images, n_views, res, n_classes, seed, rank=0, world=1):
        self.idx = list(range(rank, n_images, world))
        self.n_views, self.res, self.n_classes, self.seed = n_views, res, n_classes, seed

    def __len__(self):
        return len(self.idx)

    def __getitem__(self, i):
        j = self.idx[i]
        views = synthetic.make_views(1, self.n_views, self.res, self.seed + j)
        label = torch.tensor(j % self.n_classes)
        return views, label


# TPT/data/datautils.py:22-39
ID_to_DIRNAME = {"I": "ImageNet", "A": "imagenet-a", "K": "ImageNet-Sketch", "R": "imagenet-r",
                 "V": "imagenetv2-matched-frequency-format-val", "flower102": "oxford_flowers", "dtd": "dtd",
                 "pets": "oxford_pets", "cars": "stanford_cars", "ucf101": "ucf101", "caltech101": "caltech-101",
                 "food101": "food-101", "sun397": "sun397", "aircraft": "fgvc_aircraft", "eurosat": "eurosat",
                 "C": "imagenet-c"}


class ImageFolderViews(torch.utils.data.Dataset):
    """datasets.ImageFolder + AugMixAugmenter (tune_cls_rl.py:102-150) with the pixels left to the GPU: a sample is
    ((decoded uint8 image [H,W,3], ViewPlan), label).  The plan (random crops / flips / AugMix decisions and Pillow tap
    tables) is host work and parallelises over DataLoader workers; rank r of R sees indices r, r+R, ..."""

    def __init__(self, root, n_views, augmix, rank=0, world=1, severity=1, hard_aug=False):
        from torchvision.datasets import ImageFolder
        self.folder = ImageFolder(root)
        self.classes = self.folder.classes
        self.idx = list(range(rank, len(self.folder), world))
        self.n_views, self.augmix, self.severity, self.hard_aug = n_views, bool(augmix), severity, bool(hard_aug)

    def __len__(self):
        return len(self.idx)

    def __getitem__(self, i):
        img, label = self.folder[self.idx[i]]
        u8 = datautils._to_u8_hwc(img)
        if self.hard_aug:       # --hard_aug 1: BYOL pre-augmentation on the host (datautils.py:76-87), the rest on the GPU
            plan = datautils.sample_plan_hard(img, self.n_views, self.augmix, self.severity, host_taps=False)
        else:
            plan = datautils.sample_plan(u8.shape[1], u8.shape[0], self.n_views, self.augmix, self.severity,
                                         host_taps=False)
        return (u8, plan), torch.tensor(label)


def _collate_one(batch):
    """DataLoader collate for batch_size 1 (module-level so that worker processes can pickle it)."""
    return batch[0][0], batch[0][1].view(1)


def load_classname_table(path):
    """--classnames FILE: what the reference hard-codes in TPT/data/imagnet_prompts.py / cls_to_names.py and indexes
    with the imagenet_{a,r,v}_mask tables (tune_cls_rl.py:122-141) -- here a user-supplied file, because those tables
    are the reference's data.  Accepted: JSON {folder: name}, JSON {set_id: {folder: name} | [names]}, JSON [names]
    (index = class id), or text lines "<folder> <name>[, synonyms]" (the format of ImageNet's LOC_synset_mapping.txt)."""
    if path is None:
        return None
    with open(path) as f:
        text = f.read()
    try:
        return json.loads(text)
    except ValueError:
        table = {}
        for line in text.splitlines():
            if line.strip():
                folder, _, name = line.strip().partition(" ")
                table[folder] = name.split(",")[0].strip()
        return table


def class_names_for(folders, table, set_id):
    """Class names in ImageFolder label order for one test set.  Folder names that are WordNet ids (ImageNet, -A, -R,
    -Sketch) or integers (ImageNetV2) carry no meaning: without a table they are an error, never a prompt."""
    import re
    if isinstance(table, dict) and set_id in table and isinstance(table[set_id], (dict, list)):
        table = table[set_id]
    opaque = all(re.fullmatch(r"n\d{8}|\d+", c) for c in folders)
    if table is None:
        if opaque:
            raise SystemExit(f"test set {set_id}: folder names such as '{folders[0]}' are ids, not class names; pass "
                             "--classnames FILE (JSON {folder: name} / [names], or LOC_synset_mapping.txt lines)")
        return [c.replace("_", " ") for c in folders]
    if isinstance(table, dict):
        missing = [c for c in folders if c not in table]
        if missing and opaque:
            raise SystemExit(f"test set {set_id}: --classnames has no entry for folders {missing[:5]} ...")
        # a table of ids (e.g. wnid -> name) does not have to know the folders of sets that are named in clear
        return [str(table[c]) if c in table else c.replace("_", " ") for c in folders]
    if all(c.isdigit() for c in folders):                    # ImageNetV2: folder k holds class id k
        if max(int(c) for c in folders) >= len(table):
            raise SystemExit(f"test set {set_id}: --classnames lists {len(table)} names, folders go beyond that")
        return [str(table[int(c)]) for c in folders]
    if len(table) != len(folders):
        raise SystemExit(f"test set {set_id}: --classnames lists {len(table)} names for {len(folders)} class folders")
    return [str(t) for t in table]


def _check_supported_flags(args):
    """Reference flags this driver accepts for command-line compatibility but does not implement must not be dropped
    silently (params.py:23-73)."""
    if getattr(args, "confidence_gap", 0):
        raise NotImplementedError("--confidence_gap is an experimental reference feature that is not implemented")
    if getattr(args, "multiple_reward_models", 0) and "," not in str(args.reward_arch):
        raise NotImplementedError("--multiple_reward_models 1 needs --reward_arch 'ViT-L/14,ViT-B/16,...' (the "
                                  "reference's hard-coded list includes RN50x64, a ResNet, which is out of scope)")
    if not getattr(args, "tpt", False) and not getattr(args, "synthetic", False):
        raise NotImplementedError("without --tpt the reference runs plain batched zero-shot evaluation "
                                  "(tune_cls_rl.py:112-119); this driver implements the --tpt adaptation loop only")
    if "RN" in str(args.arch):
        raise NotImplementedError("ModifiedResNet policies are out of scope; use a ViT (-a ViT-B/16)")


def _real_dataset_root(args, set_id):
    if set_id == "I":
        return os.path.join(args.data, ID_to_DIRNAME[set_id], "val")
    return os.path.join(args.data, ID_to_DIRNAME.get(set_id.lower() if len(set_id) > 1 else set_id, set_id))


def main_worker(gpu, args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        gpu = int(os.environ.get("LOCAL_RANK", gpu))
        torch.cuda.set_device(gpu)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", gpu))
    args.gpu = gpu
    set_random_seed(args.seed)
    if not torch.cuda.is_available():
        raise SystemExit("rlcf_b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(gpu)
    device = torch.device("cuda", gpu)
    _check_supported_flags(args)
    real_data = args.data is not None and not args.synthetic
    if not real_data and not args.synthetic:
        raise SystemExit("give a dataset root (DIR, ImageFolder layout as in TPT/data/datautils.py) or --synthetic")
    synth_w = args.synthetic or getattr(args, "synthetic_weights", False)
    arch = "synthetic:" + args.arch + ":0" if synth_w else args.arch
    reward_arch = "synthetic:" + args.reward_arch + ":1" if synth_w else args.reward_arch
    datasets_, set_classnames = {}, {}
    if real_data:   # every test set has its own label space (tune_cls_rl.py:120-141)
        table = load_classname_table(getattr(args, "classnames", None))
        if getattr(args, "resolution", 224) != 224:
            raise NotImplementedError("--resolution other than 224 is not implemented (the view kernels crop to 224)")
        for set_id in args.test_sets.split("/"):
            datasets_[set_id] = ImageFolderViews(_real_dataset_root(args, set_id), args.batch_size - 1,
                                                 augmix=len(set_id) > 1, rank=rank, world=world,   # tune_cls_rl.py:109-110
                                                 hard_aug=bool(getattr(args, "hard_aug", 0)))
            set_classnames[set_id] = class_names_for(datasets_[set_id].classes, table, set_id)
        classnames = set_classnames[args.test_sets.split("/")[0]]
        args.n_classes = len(classnames)
    else:
        classnames = [f"class {i}" for i in range(args.n_classes)]
    key = arch.split(":")[1] if synth_w else args.arch
    vocab = synthetic.ARCHS[key][6] if key in synthetic.ARCHS else 49408
    tokens = synthetic.make_tokens(args.n_classes, vocab) if args.synthetic else None
    model = CLIPCLS_TTA(device, classnames, arch=arch, prompt_prefix=args.ctx_init or "a photo of a",
                        only_visual=True, momentum_update=args.momentum_update, update_freq=args.update_freq,
                        update_w=args.update_w, momentum=args.tta_momentum, only_norm=args.tune_norm,
                        tokenized_prompts=tokens)
    optimizer = torch.optim.AdamW(model.parameters(), args.lr, weight_decay=args.weight_decay)
    optim_state = deepcopy(optimizer.state_dict())
    args.reward_arch = reward_arch
    reward_model = get_reward_model(device, args)
    if args.synthetic:
        reward_model.set_class_features(tokenized_classes=synthetic.make_tokens(
            args.n_classes, synthetic.ARCHS[reward_arch.split(":")[1]][6]).to(device))
    else:
        reward_model.set_class_features(classnames=[f"{args.ctx_init or 'a photo of a'} {c}." for c in classnames])
    scaler = torch.cuda.amp.GradScaler(init_scale=1000)
    results = {}
    for set_id in args.test_sets.split("/"):
        t0 = time.time()
        if real_data:
            ds = datasets_[set_id]
            args.n_images = len(ds.folder)
            # this set's label space: new class prompts for the policy and the reward model (tune_cls_rl.py:139-143)
            args.n_classes = len(set_classnames[set_id])
            model.reset_classnames_and_state(set_classnames[set_id], args.arch)
            reward_model.set_class_features(tokenized_classes=model.tokenized_prompts)
            loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=args.workers,
                                                 collate_fn=_collate_one)
        else:
            ds = SyntheticViews(args.n_images, args.batch_size, args.resolution, args.n_classes, args.seed + 1000,
                                rank, world)
            loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=0,
                                                 collate_fn=_collate_one)
        results[set_id] = test_time_adapt_eval(loader, model, optimizer, optim_state, scaler, args, device=device,
                                               reward_model=reward_model)
        if rank == 0:
            dt = time.time() - t0
            print(f"=> Acc. on testset [{set_id}]: @1 {results[set_id][0]} / @5 {results[set_id][1]}  "
                  f"({args.n_images / dt:.1f} images/s incl. view generation)")
    if rank == 0:
        with open(os.path.join(args.output, "results.json"), "a+") as fp:
            json.dump(results, fp, indent=4)
    if world > 1:
        torch.distributed.destroy_process_group()
    return results


if __name__ == "__main__":
    _args = get_args()
    main_worker(_args.gpu, _args)
