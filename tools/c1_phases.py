#!/usr/bin/env python3
"""GPU time of the two phases of B = 1 synthesis (CUDA events around the library calls) and the host gaps.
Usage: python tools/c1_phases.py [opt=val,...]"""
import ctypes
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import efficient_tts_b200 as E  # noqa: E402
from efficient_tts_b200 import workloads as wl, _lib  # noqa: E402
from efficient_tts_b200.engine import _ptr  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    m0 = E.EfficientTTSCNN(**wl.MODEL_KWARGS).eval()
    state = {k: v.clone() for k, v in m0.state_dict().items()}
    m = E.EfficientTTSCNN(**wl.MODEL_KWARGS)
    m.load_state_dict(wl.c1_weights_patch(state))
    m = m.eval().to(dev)
    eng = m._get_engine()
    for kv in (sys.argv[1].split(",") if len(sys.argv) > 1 else []):      # opt=val,... applied to the context
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    n_iter = int(os.environ.get("ITERS", "30"))
    txt = wl.make_inference_inputs(0, 64).to(dev)
    T1 = 64
    out = {}
    for name, stack in (("stack", 1), ("per_layer", 0)):
        eng.set_option("stack", stack)
        mel, _ = eng.inference(txt)
        t2 = mel.shape[1]
        ws, n = eng.workspace_for(1, T1, t2)
        t2_dev = torch.empty(2, dtype=torch.int32, device=dev)
        ra = torch.empty(1, T1, t2, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = dict(p1_gpu=0.0, p2_gpu=0.0, p1_host=0.0, p2_host=0.0, readback=0.0)
        for i in range(n_iter + 3):
            torch.cuda.synchronize()
            ev[0].record()
            t0 = time.perf_counter()
            _lib.check(eng.lib.efts_inference_phase1(eng._h, _ptr(txt), T1, _ptr(t2_dev), _ptr(ws), ws.numel(), eng._stream()))
            t1 = time.perf_counter()
            ev[1].record()
            h = t2_dev.cpu()
            t2h = time.perf_counter()
            ev[2].record()
            _lib.check(eng.lib.efts_inference_phase2(eng._h, T1, t2, _ptr(mel), _ptr(ra), _ptr(ws), ws.numel(), eng._stream()))
            t3 = time.perf_counter()
            ev[3].record()
            torch.cuda.synchronize()
            if i >= 3:
                acc["p1_gpu"] += ev[0].elapsed_time(ev[1]); acc["p2_gpu"] += ev[2].elapsed_time(ev[3])
                acc["p1_host"] += (t1 - t0) * 1e3; acc["p2_host"] += (t3 - t2h) * 1e3; acc["readback"] += (t2h - t1) * 1e3
        out[name] = {k: round(v / n_iter, 4) for k, v in acc.items()}
    eng.set_option("stack", 1)
    print("C1PH " + json.dumps(out))
    # phase stamps of the resident kernel (SM cycles of CTA 0)
    eng.set_option("stack_trace", 1)
    buf = (ctypes.c_int64 * 64)()
    for phase in (1, 2):
        t2_dev = torch.empty(2, dtype=torch.int32, device=dev)
        _lib.check(eng.lib.efts_inference_phase1(eng._h, _ptr(txt), T1, _ptr(t2_dev), _ptr(ws), ws.numel(), eng._stream()))
        if phase == 2:
            _lib.check(eng.lib.efts_inference_phase2(eng._h, T1, t2, _ptr(mel), _ptr(ra), _ptr(ws), ws.numel(), eng._stream()))
        torch.cuda.synchronize()
        npts = (3 + 4 * 8) if phase == 1 else (1 + 4 * 7)
        _lib.check(eng.lib.efts_profile_stack_trace(eng._h, buf, npts))
        st = [buf[i] - buf[0] for i in range(npts)]
        print("TRACE phase %d cycles since start: %s" % (phase, st))
        print("TRACE phase %d deltas: %s" % (phase, [st[i + 1] - st[i] for i in range(npts - 1)]))
        _lib.check(eng.lib.efts_profile_stack_trace(eng._h, buf, 64))
        base = 3 + 4 if phase == 1 else 1 + 4     # index of "layer 1 starts" (= layer 0's second barrier passed)
        # (thread 0's phase stamps are clock reads the compiler may hoist above the CTA barrier in front of them: "gemm
        # phase done" is when the producer warp finished, the drain warps' own stamps say when the partials were stored)
        print("TRACE phase %d layer 1 roles (cycles since layer start): first operands %d, last commit %d, acc ready %d, "
              "tmem released %d, stores issued %d | gemm phase done %d | reduce loop entered %d, reduce phase done %d" % (
                  phase, *[buf[48 + k] - buf[base - 1 + 0] for k in range(5)], buf[base] - buf[base - 1],
                  buf[53] - buf[base - 1], buf[base + 2] - buf[base - 1]))
        print("TRACE phase %d layer 1 MMA steps (operands landed, cycles since layer start): %s" % (
            phase, [buf[36 + k] - buf[base - 1] for k in range(10)]))
    eng.set_option("stack_trace", 0)


if __name__ == "__main__":
    main()
