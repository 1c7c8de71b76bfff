#!/usr/bin/env python3
"""Where does the B = 1 synthesis latency go?  Wall time per inference() against the sum of its kernel times
(library profile hooks) and the host cost of enqueueing each phase."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import efficient_tts_b200 as E  # noqa: E402
from efficient_tts_b200 import workloads as wl  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    m0 = E.EfficientTTSCNN(**wl.MODEL_KWARGS).eval()
    state = {k: v.clone() for k, v in m0.state_dict().items()}
    m = E.EfficientTTSCNN(**wl.MODEL_KWARGS)
    m.load_state_dict(wl.c1_weights_patch(state))
    m = m.eval().to(dev)
    eng = m._get_engine()
    for kv in (sys.argv[1] if len(sys.argv) > 1 else "").split(","):
        if kv:
            eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    txt = wl.make_inference_inputs(0, 64).to(dev)
    for _ in range(5):
        mel, _ = m.inference(txt)
    torch.cuda.synchronize()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        mel, _ = m.inference(txt)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    names = ["text_conv", "mel_conv", "dec_conv", "linear", "energy_gemm", "softmax_expect", "imv_scan",
             "aligned_pos", "reconstruct", "expand_gemm", "duration", "loss", "embed_split"]
    eng.profile_enable(0x1FFF)
    k = 10
    for _ in range(k):
        mel, _ = m.inference(txt)
    torch.cuda.synchronize()
    prof = {names[t]: round(eng.profile_read(t)[0] / k, 4) for t in range(13)}
    cnt = {names[t]: eng.profile_read(t)[1] // k for t in range(13)}
    eng.profile_enable(0)
    print("C1 " + json.dumps(dict(T2=int(mel.shape[1]), wall_ms=wall, kernel_ms=prof, kernel_sum_ms=sum(prof.values()),
                                  launches=cnt)))


if __name__ == "__main__":
    main()
