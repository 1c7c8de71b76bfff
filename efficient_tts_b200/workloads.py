"""Seeded synthetic inputs of the BASELINE.json configs (shared by tests/ and bench.py).

Shapes follow the reference's collate output (datasets/taco2_data.py:95-139): ``text`` int64
[B, T1] padded with id 0, ``speech`` float32 [B, T2, 80] padded with zeros, and the padded dims
equal ``max(lengths)`` (utils/nets_utils.py:148).  Recipes C1..C5: SURVEY.md 8d.
"""
import numpy as np
import torch

NUM_SYMBOLS = 76
ODIM = 80
MODEL_KWARGS = dict(num_symbols=NUM_SYMBOLS, dropout_rate=0.0, use_masking=True,
                    use_weighted_masking=False, sigma=0.01)   # egs/lj/conf/...v1.yaml:17-22


def make_forward_inputs(seed, t1_lens, t2_lens):
    t1_lens = [int(x) for x in t1_lens]
    t2_lens = [int(x) for x in t2_lens]
    B, T1, T2 = len(t1_lens), max(t1_lens), max(t2_lens)
    g = torch.Generator().manual_seed(int(seed))
    text = torch.randint(0, NUM_SYMBOLS, (B, T1), generator=g)
    speech = torch.randn(B, T2, ODIM, generator=g) * 1.5 - 4.0        # log-mel-like
    tl = torch.tensor(t1_lens, dtype=torch.int64)
    sl = torch.tensor(t2_lens, dtype=torch.int64)
    text = text * (torch.arange(T1).unsqueeze(0) < tl.unsqueeze(1))
    speech = speech * (torch.arange(T2).unsqueeze(0) < sl.unsqueeze(1)).unsqueeze(-1)
    return text, tl, speech, sl


def make_inference_inputs(seed, t1):
    g = torch.Generator().manual_seed(int(seed))
    return torch.randint(0, NUM_SYMBOLS, (1, int(t1)), generator=g)


def config_lengths(name, seed=0, batch=None):
    """(t1_lens, t2_lens) of the named BASELINE.json config.  ``batch`` overrides B (bounded CPU
    samples of the same length distribution); the maxima are pinned so padded dims stay
    (200, 1200) for C3 -- the reference requires padded dim == max(lengths)."""
    if name == "C2":
        return [100] * (batch or 16), [800] * (batch or 16)
    if name == "C3":
        t1 = np.random.default_rng(seed).integers(50, 201, batch or 256)
        if seed == 0 and batch is None:
            assert int(t1.sum()) == 32825
        if int(t1.max()) != 200:
            t1[int(t1.argmax())] = 200
        return t1.tolist(), (6 * t1).tolist()
    if name == "C5":
        return [300] * (batch or 32), [2000] * (batch or 32)
    raise KeyError(name)


def c1_weights_patch(state_dict):
    """C1 recipe: scale the duration head so a 64-phoneme input yields ~512 frames."""
    sd = dict(state_dict)
    sd["duration_predictor.linear.weight"] = sd["duration_predictor.linear.weight"] * 0.05
    sd["duration_predictor.linear.bias"] = torch.full_like(sd["duration_predictor.linear.bias"],
                                                           float(np.log(9.0)))
    return sd


# ---------------------------------------------------------------- HiFi-GAN V1 generator
HIFIGAN_V1 = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                  upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                  resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=80)
# vocoders/HiFiGAN_LJ_V1/config.json


class AttrDict(dict):
    """Attribute-style config like vocoders/env.py ``AttrDict``."""
    __getattr__ = dict.__getitem__


def vocoder_state_dict(seed=4321, h=HIFIGAN_V1):
    """Seeded generator weights with unit-scale layers (the reference's own init, N(0, 0.01), makes a random
    generator's output vanish).  The test suite's CPU checker uses the same recipe (a CPU test keeps the two equal);
    it is restated here because the product package and bench.py's CUDA arm never import test infrastructure."""
    import math
    g = torch.Generator().manual_seed(seed)
    w = {}

    def conv(prefix, cout, cin, k, transposed=False, g_layer=1.0):
        b = g_layer / math.sqrt(cin * k)
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        v = (torch.rand(*shape, generator=g) * 2 - 1) * b
        norm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
        w[prefix + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * b
        w[prefix + ".weight_g"] = norm * (1.0 + 0.05 * torch.randn(norm.shape, generator=g))
        w[prefix + ".weight_v"] = v

    c0 = h["upsample_initial_channel"]
    conv("conv_pre", c0, h["num_mels"], 7)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        conv("ups.%d" % i, c0 // 2 ** (i + 1), c0 // 2 ** i, k, transposed=True, g_layer=1.5 * math.sqrt(u))
    for i in range(len(h["upsample_rates"])):
        ch = c0 // 2 ** (i + 1)
        for j, (k, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            for m in range(len(dil)):
                conv("resblocks.%d.convs1.%d" % (i * nk + j, m), ch, ch, k)
            for m in range(len(dil)):
                conv("resblocks.%d.convs2.%d" % (i * nk + j, m), ch, ch, k)
    conv("conv_post", 1, c0 // 2 ** len(h["upsample_rates"]), 7, g_layer=4.0)
    return w


def make_mel(seed, batch, frames, num_mels=80):
    """Log-mel-like synthetic vocoder input [B, num_mels, T]."""
    g = torch.Generator().manual_seed(int(seed))
    return torch.randn(batch, num_mels, frames, generator=g) * 1.5 - 4.0
