#!/usr/bin/env python3
"""B = 1 synthesis latency (config C1), resident stack against one launch per layer; per-phase GPU time from CUDA
events around the two library calls.  Usage: python tools/c1_latency.py [opt=val,...]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import efficient_tts_b200 as E  # noqa: E402
from efficient_tts_b200 import workloads as wl  # noqa: E402


def timed(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    m0 = E.EfficientTTSCNN(**wl.MODEL_KWARGS).eval()
    state = {k: v.clone() for k, v in m0.state_dict().items()}
    m = E.EfficientTTSCNN(**wl.MODEL_KWARGS)
    m.load_state_dict(wl.c1_weights_patch(state))
    m = m.eval().to(dev)
    eng = m._get_engine()
    txt = wl.make_inference_inputs(0, 64).to(dev)
    res = {}
    for name, opts in (("stack", dict(stack=1)), ("per_layer", dict(stack=0))):
        for k, v in opts.items():
            eng.set_option(k, v)
        n0 = eng.launch_count()
        mel, _ = m.inference(txt)
        res[name] = dict(launches=eng.launch_count() - n0, T2=int(mel.shape[1]),
                         wall_ms=timed(lambda: m.inference(txt)),
                         wall_ms_no_flag_check=timed(lambda: eng.inference(txt, check_numerics=False)))
    eng.set_option("stack", 1)
    print("C1LAT " + json.dumps(res))
    if os.environ.get("PROFILE"):
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(300):
            m.inference(txt)
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(18)


if __name__ == "__main__":
    main()
