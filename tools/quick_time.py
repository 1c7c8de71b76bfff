#!/usr/bin/env python3
"""Quick device timing of forward() at a named config (bring-up helper; bench.py is the contract)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import efts_oracle as orc  # noqa: E402  (weights recipe only)
from tests.cases import config_lengths, make_forward_inputs  # noqa: E402
import efficient_tts_b200 as E  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    opts = dict(kv.split("=") for kv in sys.argv[2].split(",")) if len(sys.argv) > 2 and sys.argv[2] else {}
    dev = torch.device("cuda", 0)
    m = E.EfficientTTSCNN(num_symbols=76, dropout_rate=0.0, use_masking=True, sigma=0.01)
    m.load_state_dict(orc.make_weights(seed=1234))
    m = m.eval().to(dev)
    t1, t2 = config_lengths(name)
    text, tl, speech, sl = (x.to(dev) for x in make_forward_inputs(0, t1, t2))
    eng = m._get_engine()
    for k, v in opts.items():
        eng.set_option(k, int(v))
    for _ in range(3):
        out = m(text=text, text_lengths=tl, speech=speech, speech_lengths=sl)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 5
    ev[0].record()
    for _ in range(n):
        eng.forward(text, tl, speech, sl)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    frames = int(sl.sum())
    print("QUICK " + json.dumps(dict(config=name, opts=opts, ms=ms, valid_frames=frames,
                                     frames_per_s=frames / ms * 1e3, stats=out[1],
                                     launches=eng.launch_count())))


if __name__ == "__main__":
    main()
