#!/usr/bin/env python3
"""Golden fixtures for the HiFi-GAN V1 generator from the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``):  ``python tests/golden/make_golden_hifigan.py``

``nntts.vocoders.hifigan_model`` is imported as-is; its only missing dependency here is matplotlib, which
``vocoders/utils.py`` imports for a plotting helper the generator never calls, so an empty stand-in module
is registered before the import.  The reference ``Generator`` is built from the V1 config, the
deterministic weights of ``oracle.hifigan_oracle.make_weights`` are loaded through its own
``load_state_dict`` (strict), and ``forward`` runs on seeded mel inputs, with and without
``remove_weight_norm()``.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

for name in ("matplotlib", "matplotlib.pylab"):
    if name not in sys.modules:
        try:
            __import__(name)
        except ImportError:
            mod = types.ModuleType(name)
            mod.use = lambda *a, **k: None
            sys.modules[name] = mod
if not hasattr(sys.modules["matplotlib"], "pylab"):
    sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]

from oracle import hifigan_oracle as hor  # noqa: E402  (weights + input recipes only)
from nntts.vocoders.env import AttrDict  # noqa: E402
from nntts.vocoders.hifigan_model import Generator  # noqa: E402


def case(name, seed, batch, frames, cfg=hor.V1_CONFIG):
    w = hor.make_weights(seed=4321, h=cfg)
    gen = Generator(AttrDict(cfg))
    missing = gen.load_state_dict(w, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    gen.eval()
    mel = hor.make_mel(seed, batch, frames)
    with torch.no_grad():
        y = gen(mel)
        gen.remove_weight_norm()
        y2 = gen(mel)
    assert torch.equal(y, y2) or (y - y2).abs().max() < 1e-6
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, batch=batch, frames=frames,
                        audio=y.numpy())
    print(name, tuple(y.shape), "abs max", float(y.abs().max()), "std", float(y.std()))


if __name__ == "__main__":
    case("hifigan_small", 7, 1, 12)
    case("hifigan_batch", 8, 2, 9)
    case("hifigan_v2", 9, 2, 10, hor.V2_CONFIG)
    case("hifigan_v3", 10, 2, 11, hor.V3_CONFIG)
