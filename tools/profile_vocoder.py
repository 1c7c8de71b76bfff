#!/usr/bin/env python3
"""HiFi-GAN V1 generator on B x T frames, a few calls (for ncu launch lists).  python tools/profile_vocoder.py [B] [T] [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from efficient_tts_b200 import workloads as wl  # noqa: E402
from efficient_tts_b200.vocoder import Generator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 800
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
voc = Generator(wl.AttrDict(wl.HIFIGAN_V1))
voc.load_state_dict(wl.vocoder_state_dict())
voc = voc.eval().to(dev)
mel = wl.make_mel(1, B, T).to(dev)
for _ in range(n):
    y = voc(mel)
torch.cuda.synchronize()
print("done", tuple(y.shape))
