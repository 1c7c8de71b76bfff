#!/usr/bin/env python3
"""Timing-only experiment (results are wrong under debug_mask): how much of a conv launch is the cost of the
A / W tile traffic from L2?  debug_mask 4 = W tiles are not reloaded, 8 = A boxes are not reloaded."""
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import efficient_tts_b200 as E  # noqa: E402
from efficient_tts_b200 import workloads as wl  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    m = E.EfficientTTSCNN(**wl.MODEL_KWARGS).eval().to(dev)
    eng = m._get_engine()
    t1, t2 = wl.config_lengths("C3", seed=0)
    text, tl, speech, sl = (t.to(dev) for t in wl.make_forward_inputs(0, t1, t2))
    names = ["text_conv", "mel_conv", "dec_conv", "linear", "energy_gemm", "softmax_expect", "imv_scan",
             "aligned_pos", "reconstruct", "expand_gemm", "duration", "loss", "embed_split"]
    for opts in sys.argv[1:] or [""]:
        for kv in opts.split(","):
            if kv:
                eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        for _ in range(3):
            eng.forward(text, tl, speech, sl)
        torch.cuda.synchronize()
        clocks = []
        stop = threading.Event()

        def sample():
            while not stop.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                          "-i", "0"], capture_output=True, text=True, timeout=5).stdout.strip()
                    clocks.append(out)
                except Exception:
                    pass
                time.sleep(0.05)
        th = threading.Thread(target=sample)
        th.start()
        eng.profile_enable(0x1FFF)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n = 30
        ev[0].record()
        for _ in range(n):
            eng.forward(text, tl, speech, sl)
        ev[1].record()
        torch.cuda.synchronize()
        stop.set()
        th.join()
        prof = {names[t]: round(eng.profile_read(t)[0] / n, 4) for t in range(13)}
        eng.profile_enable(0)
        print("EXP " + json.dumps(dict(opts=opts, ms=ev[0].elapsed_time(ev[1]) / n, clocks=clocks[len(clocks) // 2:][:3],
                                       prof={k: prof[k] for k in ("text_conv", "mel_conv", "dec_conv", "linear")})), flush=True)
        eng.set_option("debug_mask", 0)


if __name__ == "__main__":
    main()
