"""CPU oracle for the HiFi-GAN V1 generator.  TEST INFRASTRUCTURE ONLY.

Restates ``nntts.vocoders.hifigan_model.Generator.forward`` (vocoders/hifigan_model.py:95-136) and
``ResBlock1.forward`` (:56-63) functionally over a plain dict of tensors with the same torch CPU ops in
the same order, so that it has no dependency on ``/root/reference`` and can travel to the GPU box.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it.

Parity pinning: the reference ships no tests for the vocoder; this file is pinned to outputs of the
unmodified reference ``Generator`` imported in the build container (``tests/golden/make_golden_hifigan.py``
-> ``tests/golden/hifigan_*.npz``; ``tests/test_oracle_golden.py`` checks them).

Reference citations are relative to ``/root/reference/nntts``.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]

LRELU_SLOPE = 0.1                       # vocoders/hifigan_model.py:11

# vocoders/HiFiGAN_LJ_V1/config.json (the generator ``bin/inference.py:85`` loads)
V1_CONFIG = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                 upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                 resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=80)


# config_v2.json / config_v3.json of the HiFi-GAN release the reference's Generator class was taken from: the same
# class builds them (V2: ResBlock1 with 128 initial channels; V3: ResBlock2, three upsamplers, dilations up to 12)
V2_CONFIG = dict(V1_CONFIG, upsample_initial_channel=128)
V3_CONFIG = dict(resblock="2", upsample_rates=[8, 8, 4], upsample_kernel_sizes=[16, 16, 8],
                 upsample_initial_channel=256, resblock_kernel_sizes=[3, 5, 7],
                 resblock_dilation_sizes=[[1, 2], [2, 6], [3, 12]], num_mels=80)


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """vocoders/utils.py ``get_padding``."""
    return int((kernel_size * dilation - dilation) / 2)


def make_weights(seed: int = 4321, h: dict = V1_CONFIG, gain: float = 1.0) -> Weights:
    """Deterministic random ``state_dict`` with the reference's key names and shapes
    (vocoders/hifigan_model.py:31-54, 97-118): every conv is weight-normed (``weight_g`` / ``weight_v``).
    The reference initialises with N(0, 0.01), which makes a random generator's output vanish; the
    fixtures use PyTorch's default scale (uniform +-1/sqrt(fan_in)) so that every layer carries signal."""
    g = torch.Generator().manual_seed(seed)
    w: Weights = {}

    def conv(prefix, cout, cin, k, transposed=False, g_layer=1.0):
        fan_in = cin * k
        b = gain * g_layer / math.sqrt(fan_in)
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        v = (torch.rand(*shape, generator=g) * 2 - 1) * b
        norm = v.flatten(1).norm(dim=1).view(-1, 1, 1)          # weight_norm dim 0 (also for ConvTranspose1d)
        w[prefix + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * b
        w[prefix + ".weight_g"] = norm * (1.0 + 0.05 * torch.randn(norm.shape, generator=g))
        w[prefix + ".weight_v"] = v

    c0 = h["upsample_initial_channel"]
    conv("conv_pre", c0, h["num_mels"], 7)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        # a transposed conv with stride u only has k / u taps under every output sample
        conv(f"ups.{i}", c0 // 2 ** (i + 1), c0 // 2 ** i, k, transposed=True, g_layer=1.5 * math.sqrt(u))
    for i in range(len(h["upsample_rates"])):
        ch = c0 // 2 ** (i + 1)
        for j, (k, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            if h["resblock"] == "1":
                for m in range(len(dil)):
                    conv(f"resblocks.{i * nk + j}.convs1.{m}", ch, ch, k)
                for m in range(len(dil)):
                    conv(f"resblocks.{i * nk + j}.convs2.{m}", ch, ch, k)
            else:                                   # ResBlock2 (:71-81): one conv per dilation
                for m in range(len(dil)):
                    conv(f"resblocks.{i * nk + j}.convs.{m}", ch, ch, k)
    conv("conv_post", 1, c0 // 2 ** len(h["upsample_rates"]), 7, g_layer=4.0)   # waveform spans tanh's range
    return w


def conv_weight(w: Weights, prefix: str) -> torch.Tensor:
    """Effective weight of a weight-normed conv (``g * v / ||v||`` over dim 0), or the folded ``.weight``
    left by ``remove_weight_norm`` (vocoders/hifigan_model.py:138-145)."""
    if prefix + ".weight" in w:
        return w[prefix + ".weight"]
    return torch._weight_norm(w[prefix + ".weight_v"], w[prefix + ".weight_g"], 0)


def resblock1(x: torch.Tensor, w: Weights, prefix: str, k: int, dilation) -> torch.Tensor:
    """vocoders/hifigan_model.py:56-63."""
    for m, d in enumerate(dilation):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, conv_weight(w, f"{prefix}.convs1.{m}"), w[f"{prefix}.convs1.{m}.bias"],
                      dilation=d, padding=get_padding(k, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, conv_weight(w, f"{prefix}.convs2.{m}"), w[f"{prefix}.convs2.{m}.bias"],
                      dilation=1, padding=get_padding(k, 1))
        x = xt + x
    return x


def resblock2(x: torch.Tensor, w: Weights, prefix: str, k: int, dilation) -> torch.Tensor:
    """vocoders/hifigan_model.py:83-88."""
    for m, d in enumerate(dilation):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, conv_weight(w, f"{prefix}.convs.{m}"), w[f"{prefix}.convs.{m}.bias"], dilation=d,
                      padding=get_padding(k, d))
        x = xt + x
    return x


def generator_forward(w: Weights, mel_bct: torch.Tensor, h: dict = V1_CONFIG) -> torch.Tensor:
    """vocoders/hifigan_model.py:120-136: mel [B, 80, T] -> waveform [B, 1, T * prod(upsample_rates)]."""
    resblock = resblock1 if h["resblock"] == "1" else resblock2          # :102
    nk = len(h["resblock_kernel_sizes"])
    x = F.conv1d(mel_bct, conv_weight(w, "conv_pre"), w["conv_pre.bias"], padding=3)            # :121
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)                                                         # :123
        x = F.conv_transpose1d(x, conv_weight(w, f"ups.{i}"), w[f"ups.{i}.bias"], stride=u,
                               padding=(k - u) // 2)                                             # :124
        xs = None
        for j in range(nk):                                                                      # :126-130
            r = resblock(x, w, f"resblocks.{i * nk + j}", h["resblock_kernel_sizes"][j],
                         h["resblock_dilation_sizes"][j])
            xs = r if xs is None else xs + r
        x = xs / nk                                                                              # :131
    x = F.leaky_relu(x)                                                                          # :132 (slope 0.01)
    x = F.conv1d(x, conv_weight(w, "conv_post"), w["conv_post.bias"], padding=3)                 # :133
    return torch.tanh(x)                                                                         # :134


def make_mel(seed: int, batch: int, frames: int, num_mels: int = 80) -> torch.Tensor:
    """Log-mel-like synthetic input [B, num_mels, T] (seeded)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, num_mels, frames, generator=g) * 1.5 - 4.0
