#!/usr/bin/env python3
"""HiFi-GAN V1 generator on B200: latency at the C1 length (B = 1, 517 frames = 6.0 s of audio) and throughput on a
batch, plus the text -> waveform RTF the reference defines (bin/inference.py:100-111: inference + vocoder)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import efficient_tts_b200 as E  # noqa: E402
from efficient_tts_b200 import workloads as wl  # noqa: E402
from efficient_tts_b200.vocoder import Generator  # noqa: E402

V1 = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
          upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
          resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])


class H(dict):
    __getattr__ = dict.__getitem__


def timed(fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def main():
    dev = torch.device("cuda", 0)
    from oracle import hifigan_oracle as hor      # weights recipe only (unit-scale layers; the default init is N(0, 0.01))
    g = Generator(H(V1))
    g.load_state_dict(hor.make_weights(seed=4321))
    opts = dict(kv.split("=") for kv in sys.argv[1].split(",")) if len(sys.argv) > 1 and sys.argv[1] else {}
    if opts:
        g.set_engine_options(**{k: int(v) for k, v in opts.items()})
    g = g.eval().to(dev)
    out = {}
    for B, T, n in ((1, 517, 20), (16, 800, 5)):
        mel = (torch.randn(B, 80, T) * 1.5 - 4.0).to(dev)
        for _ in range(3):
            y = g(mel)
        dt = timed(lambda: g(mel), n)
        audio_s = B * T * 256 / 22050.0
        out["B%d_T%d" % (B, T)] = dict(ms=dt * 1e3, samples_per_s=B * T * 256 / dt, rtf=dt / audio_s,
                                       launches=g._get_engine().launch_count())
    # text -> waveform, B = 1 (the reference's RTF)
    torch.manual_seed(1234)
    m0 = E.EfficientTTSCNN(**wl.MODEL_KWARGS).eval()
    m = E.EfficientTTSCNN(**wl.MODEL_KWARGS)
    m.load_state_dict(wl.c1_weights_patch({k: v.clone() for k, v in m0.state_dict().items()}))
    m = m.eval().to(dev)
    txt = wl.make_inference_inputs(0, 64).to(dev)

    def tts():
        mel, _ = m.inference(txt)
        return g(mel.transpose(1, 2))
    for _ in range(3):
        y = tts()
    dt = timed(tts, 20)
    out["text_to_wave_c1"] = dict(ms=dt * 1e3, samples=int(y.shape[-1]), rtf=dt / (y.shape[-1] / 22050.0))
    out["opts"] = opts
    print("VOC " + json.dumps(out))


if __name__ == "__main__":
    main()
