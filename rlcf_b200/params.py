"""Command-line flags of the TTA drivers, names and defaults as in TPT/params.py:13-98 (plus the few this
implementation adds: --images_per_step, --synthetic, --n_images, --n_classes)."""
from __future__ import annotations

import argparse
import json
import os


def none_or_str(value):
    return None if value == "None" else value


def build_parser():
    parser = argparse.ArgumentParser(description="Test-time adaptation with CLIP reward (RLCF), B200-native")
    parser.add_argument("data", metavar="DIR", nargs="?", default=None, help="path to dataset root")
    parser.add_argument("--test_sets", type=str, default="A/R/V/K/I")
    parser.add_argument("--dataset_mode", type=str, default="test")
    parser.add_argument("--output", type=str, default="exp_01")
    parser.add_argument("-a", "--arch", metavar="ARCH", default="RN50")
    parser.add_argument("--resolution", default=224, type=int)
    parser.add_argument("--workers", default=8, type=int)
    parser.add_argument("--batch_size", default=64, type=int, help="number of views per test image")
    parser.add_argument("--lr", "--learning-rate", default=5e-3, type=float, dest="lr")
    parser.add_argument("--weight_decay", default=5e-4, type=float)
    parser.add_argument("-p", "--print-freq", default=500, type=int)
    parser.add_argument("--gpu", default=0, type=int)
    parser.add_argument("--tpt", action="store_true", default=False)
    parser.add_argument("--selection_p", default=0.1, type=float)
    parser.add_argument("--tta_steps", default=1, type=int)
    parser.add_argument("--n_ctx", default=4, type=int)
    parser.add_argument("--ctx_init", default=None, type=str)
    parser.add_argument("--cocoop", action="store_true", default=False)
    parser.add_argument("--load", default=None, type=none_or_str)
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--hard_aug", type=int, default=0)
    parser.add_argument("--augmix", type=int, default=1)
    # RL config
    parser.add_argument("--sample_k", type=int, default=5)
    parser.add_argument("--multiple_reward_models", type=int, default=0)
    parser.add_argument("--reward_arch", type=str, default="ViT-L/14")
    parser.add_argument("--reward_process", type=int, default=1)
    parser.add_argument("--process_batch", type=int, default=0)
    parser.add_argument("--reward_amplify", type=int, default=0)
    parser.add_argument("--weighted_scores", type=int, default=1)
    parser.add_argument("--confidence_gap", type=int, default=0)
    parser.add_argument("--confidence_gap_w", type=float, default=0.5)
    parser.add_argument("--min_entropy_reg", type=int, default=0)
    parser.add_argument("--min_entropy_w", type=float, default=0.1)
    parser.add_argument("--momentum_update", type=int, default=0)
    parser.add_argument("--update_freq", type=int, default=256)
    parser.add_argument("--update_w", type=float, default=1.0)
    parser.add_argument("--tta_momentum", type=float, default=0.9999)
    parser.add_argument("--tune_norm", type=int, default=0)
    parser.add_argument("--prior_strength", type=int, default=-1)
    parser.add_argument("--corruption", type=str, default="defocus_blur")
    parser.add_argument("--level", type=str, default="5")
    parser.add_argument("--kd_loss", type=str, default="KD", choices=["KD", "DKD", "ATKD"])
    # additions of this implementation
    parser.add_argument("--images_per_step", type=int, default=8, help="independent test images adapted per launch")
    parser.add_argument("--synthetic", action="store_true", help="synthetic weights / views / labels (offline)")
    parser.add_argument("--synthetic_weights", action="store_true",
                        help="real images from DIR (views generated on the GPU) with seeded random-init CLIP weights")
    parser.add_argument("--classnames", type=str, default=None,
                        help="class-name table for test sets whose folders are ids (ImageNet wnids, ImageNetV2 "
                             "integers): JSON {folder: name} / {set_id: ...} / [names] or LOC_synset_mapping.txt lines")
    parser.add_argument("--n_images", type=int, default=64, help="synthetic: test images per dataset")
    parser.add_argument("--n_classes", type=int, default=200, help="synthetic: classes per dataset")
    return parser


def get_args(argv=None):
    args = build_parser().parse_args(argv)
    os.makedirs(args.output, exist_ok=True)
    assert args.gpu is not None
    save_hp_to_json(args.output, args)
    return args


def save_hp_to_json(directory, args):
    with open(os.path.join(directory, "hparams_train.json"), "w") as f:
        json.dump(vars(args), f, indent=4, sort_keys=True)
