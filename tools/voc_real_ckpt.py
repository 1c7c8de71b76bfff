#!/usr/bin/env python3
"""One-off check of the generator path with a TRAINED checkpoint (weights with the scales training leaves, not the
seeded unit-scale recipe of the tests).  The checkpoint is not part of this repository: pass its path, e.g. the
reference's vocoders/HiFiGAN_LJ_V1/generator_v1 copied next to the snapshot for one GPU session.  Compares the CUDA
path with the CPU oracle on the same weights and seeded log-mel-like inputs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hifigan_oracle as hor  # noqa: E402
from efficient_tts_b200 import workloads as wl  # noqa: E402
from efficient_tts_b200.vocoder import Generator  # noqa: E402


def main():
    path = sys.argv[1]
    sd = torch.load(path, map_location="cpu")["generator"]
    sd = {k: v.float() for k, v in sd.items()}
    dev = torch.device("cuda", 0)
    g = Generator(wl.AttrDict(wl.HIFIGAN_V1))
    res = g.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    g = g.eval().to(dev)
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    for B, T in ((1, 100), (2, 517)):
        mel = hor.make_mel(1000 + T, B, T)
        with torch.no_grad():
            ref = hor.generator_forward(sd, mel)
        y = g(mel.to(dev)).cpu()
        d = (y - ref).abs()
        out["B%d_T%d" % (B, T)] = dict(max_abs=float(d.max()), rms=float(d.pow(2).mean().sqrt()),
                                       signal_max=float(ref.abs().max()), signal_rms=float(ref.pow(2).mean().sqrt()))
    g.remove_weight_norm()
    y2 = g(mel.to(dev)).cpu()
    out["after_remove_weight_norm_max_abs"] = float((y2 - ref).abs().max())
    print("REALCKPT " + json.dumps(out))


if __name__ == "__main__":
    main()
