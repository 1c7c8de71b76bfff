#!/usr/bin/env python3
"""Robustness check of the resident layer-stack kernels: 400 random text lengths (3 - 200 tokens; predicted lengths from a
few dozen to 1 600 frames, i.e. both the stack and its per-layer fallback), resident stack vs one launch per layer compared
bitwise, then 3 000 back-to-back synthesis calls.  python tools/stress_stack.py"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import efficient_tts_b200 as E
from efficient_tts_b200 import workloads as wl
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
m0 = E.EfficientTTSCNN(**wl.MODEL_KWARGS).eval()
state = {k: v.clone() for k, v in m0.state_dict().items()}
m = E.EfficientTTSCNN(**wl.MODEL_KWARGS); m.load_state_dict(wl.c1_weights_patch(state)); m = m.eval().to(dev)
eng = m._get_engine()
g = torch.Generator().manual_seed(0)
n_bad = 0; t0 = time.time(); shapes = set()
for it in range(400):
    T1 = int(torch.randint(3, 201, (1,), generator=g))
    txt = wl.make_inference_inputs(it, T1).to(dev)
    eng.set_option("stack", 1)
    try:
        mel, ra = m.inference(txt)
    except RuntimeError as ex:
        if "predicted length" in str(ex): continue
        raise
    eng.set_option("stack", 0)
    mel2, ra2 = m.inference(txt)
    shapes.add((T1, mel.shape[1]))
    if not (torch.equal(mel, mel2) and torch.equal(ra, ra2)):
        n_bad += 1
        print("MISMATCH at T1", T1, "T2", mel.shape[1], float((mel - mel2).abs().max()))
eng.set_option("stack", 1)
# back-to-back calls without host syncs in between other than the library's own
txt = wl.make_inference_inputs(0, 64).to(dev)
ref, _ = m.inference(txt)
for it in range(3000):
    mel, _ = eng.inference(txt, check_numerics=False)
torch.cuda.synchronize()
print("STRESS shapes %d, mismatches %d, T2 range %d..%d, repeat-equal %s, %.1f s" % (
    len(shapes), n_bad, min(s[1] for s in shapes), max(s[1] for s in shapes), bool(torch.equal(mel, ref)), time.time() - t0))
