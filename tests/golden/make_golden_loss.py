#!/usr/bin/env python3
"""Golden fixture for the criterion: the UNMODIFIED ``nntts.losses.fastspeech_loss.FastSpeechLoss`` (forward values and
torch-autograd gradients) on seeded inputs.  Run in the build container only (needs /root/reference):
``python tests/golden/make_golden_loss.py`` -> ``tests/golden/loss_cases.npz``."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from nntts.losses.fastspeech_loss import FastSpeechLoss  # noqa: E402
from oracle.efts_oracle import make_loss_inputs  # noqa: E402

CASES = [  # name, seed, B, T1, T2, odim, use_masking, use_mse
    ("mask_mse", 0, 3, 11, 37, 80, True, True),          # the model's configuration (models/efficient_tts.py:115-117)
    ("nomask_mse", 1, 2, 7, 20, 80, False, True),
    ("mask_l1", 2, 4, 9, 25, 8, True, False),
]


def main():
    out = {}
    for name, seed, B, T1, T2, odim, use_masking, use_mse in CASES:
        mel, dur, ys, ds, il, ol = make_loss_inputs(seed, B, T1, T2, odim)
        mel.requires_grad_(True)
        dur.requires_grad_(True)
        crit = FastSpeechLoss(use_masking=use_masking, use_weighted_masking=False, use_mse=use_mse)
        ml, dl = crit(None, mel, dur, ys, ds, il, ol)
        (ml * 0.7 + dl * 1.3).backward()
        out[name + ".cfg"] = np.array([seed, B, T1, T2, odim, int(use_masking), int(use_mse)])
        out[name + ".losses"] = np.array([ml.item(), dl.item()], dtype=np.float64)
        out[name + ".grad_mel"] = mel.grad.numpy()
        out[name + ".grad_dur"] = dur.grad.numpy()
        print(name, float(ml), float(dl))
    np.savez_compressed(os.path.join(HERE, "loss_cases.npz"), **out)


if __name__ == "__main__":
    main()
