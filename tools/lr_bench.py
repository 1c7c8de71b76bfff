#!/usr/bin/env python3
"""Length-regulator throughput at the C3 shape (B = 256, 50-200 tokens, D = 512, ~6 frames per token):
algorithmic bytes (4*D*(sum T1_b + sum Tout_b) + 8*B*T1 + 8*sum Tout index) over the time of
plan + forward kernels, measured with CUDA events."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from efficient_tts_b200 import engine as eng  # noqa: E402
from efficient_tts_b200 import workloads as wl  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    t1, _ = wl.config_lengths("C3", seed=0)
    B, T1, D = len(t1), max(t1), 512
    g = torch.Generator().manual_seed(3)
    xs = torch.randn(B, T1, D, generator=g).to(dev)
    ds = torch.randint(3, 10, (B, T1), generator=g)
    il = torch.tensor(t1, dtype=torch.int64)
    for b in range(B):
        ds[b, t1[b]:] = 0
    ds, il = ds.to(dev), il.to(dev)
    for _ in range(3):
        out, idx = eng.length_regulator(xs, ds, il, return_index=True)
    torch.cuda.synchronize()
    n = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        out, idx = eng.length_regulator(xs, ds, il, return_index=True)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    tout = int(ds.sum())
    by = 4 * D * (int(il.sum()) + tout) + 8 * B * T1 + 8 * tout
    written = 4 * D * B * out.shape[1]
    peak = 6550.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print("LR " + json.dumps(dict(B=B, T1=T1, D=D, Tout_max=int(out.shape[1]), frames=tout, ms_per_call=ms,
                                  algorithmic_bytes=by, padded_output_bytes=written,
                                  achieved_gbs=by / ms / 1e6, achieved_gbs_incl_padding=(by - 4 * D * tout + written) / ms / 1e6,
                                  peak_gbs=peak, note="ms includes the plan kernel, the host read-back of Tout and the "
                                  "output allocation (the reference's own contract); the forward kernel alone is in the ncu capture")))


if __name__ == "__main__":
    main()
