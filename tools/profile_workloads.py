#!/usr/bin/env python3
"""Small fixed workloads for ncu (tools/gpu_profile.sh): a few calls of one path each, no timing, no CPU work.

  python tools/profile_workloads.py c3 | c1 | lr | frontend | train
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import efficient_tts_b200 as E  # noqa: E402
from efficient_tts_b200 import workloads as wl  # noqa: E402


def model(dev, c1=False):
    torch.manual_seed(1234)
    m = E.EfficientTTSCNN(**wl.MODEL_KWARGS).eval()
    if c1:
        sd = wl.c1_weights_patch({k: v.clone() for k, v in m.state_dict().items()})
        m = E.EfficientTTSCNN(**wl.MODEL_KWARGS)
        m.load_state_dict(sd)
        m = m.eval()
    return m.to(dev)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "c3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda", 0)
    if what == "c3":
        m = model(dev)
        t1, t2 = wl.config_lengths("C3")
        a = [t.to(dev) for t in wl.make_forward_inputs(0, t1, t2)]
        for _ in range(n):
            m._get_engine().forward(*a)
    elif what == "c1":
        m = model(dev, c1=True)
        txt = wl.make_inference_inputs(0, 64).to(dev)
        for _ in range(n):
            m.inference(txt)
    elif what == "lr":
        from efficient_tts_b200.engine import length_regulator
        t1, _ = wl.config_lengths("C3")
        g = torch.Generator().manual_seed(3)
        xs = torch.randn(256, 200, 512, generator=g).to(dev)
        ds = torch.randint(3, 10, (256, 200), generator=g)
        for b in range(256):
            ds[b, t1[b]:] = 0
        for _ in range(n):
            length_regulator(xs, ds.to(dev), torch.tensor(t1).to(dev))
    elif what == "frontend":
        from efficient_tts_b200.frontend import LogMelFrontend
        _, t2 = wl.config_lengths("C3")
        lens = torch.tensor([f * 256 for f in t2])
        audio = ((torch.rand(256, int(lens.max())) - 0.5) * (torch.arange(int(lens.max()))[None] < lens[:, None])).to(dev)
        fe = LogMelFrontend(dev)
        for _ in range(n):
            fe(audio, lens.to(dev))
    elif what == "train":
        from efficient_tts_b200.engine import train_context
        tc = train_context(dev)
        g = torch.Generator().manual_seed(12)
        L, B, T, C, k = 2, 256, 1200, 512, 5
        x = torch.randn(B, T, C, generator=g).to(dev)
        w = (torch.randn(L, C, C, k, generator=g) / np.sqrt(C * k)).to(dev)
        b = (torch.randn(L, C, generator=g) * 0.1).to(dev)
        for _ in range(n):
            acts, us = tc.resconv_fwd(x, w, b)
            tc.resconv_bwd(x, acts, us, w)
    torch.cuda.synchronize()
    print("done", what)


if __name__ == "__main__":
    main()
