#!/usr/bin/env python3
"""Front-end precision: device log-mel vs the fp32 reference restatement and vs a float64 evaluation, per chunk_kb."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend_oracle as fo  # noqa: E402
from efficient_tts_b200.frontend import LogMelFrontend  # noqa: E402


def mel64(y):
    mel = torch.from_numpy(fo.slaney_mel_basis()).double()
    w = torch.hann_window(1024, dtype=torch.float64)
    y = torch.nn.functional.pad(y.double().unsqueeze(1), (384, 384), mode="reflect").squeeze(1)
    s = torch.stft(y, 1024, hop_length=256, win_length=1024, window=w, center=False, return_complex=True)
    s = torch.sqrt(torch.view_as_real(s).pow(2).sum(-1) + 1e-9)
    return torch.log(torch.clamp(mel @ s, min=1e-5))


def main():
    dev = torch.device("cuda", 0)
    lengths = [22050, 9000, 10240, 30000, 50000]
    audio = fo.make_audio(7, lengths)
    for ckb in (2, 1):
        fe = LogMelFrontend(dev)
        fe.set_option("chunk_kb", ckb)
        worst = dict(o32=0, o64=0, r64=0)
        for b, L in enumerate(lengths):
            y = audio[b:b + 1, :L]
            ours = fe(y.to(dev))[0][0].double().cpu().T
            r32 = fo.mel_spectrogram(y)[0].double()
            r64 = mel64(y)[0]
            peak = r64.exp().max(0, keepdim=True).values
            strong = r64.exp() >= 1e-3 * peak
            e = dict(o32=(ours - r32).abs()[strong].max().item(), o64=(ours - r64).abs()[strong].max().item(),
                     r64=(r32 - r64).abs()[strong].max().item())
            lin = dict(o64=((ours.exp() - r64.exp()).abs() / peak).max().item(), r64=((r32.exp() - r64.exp()).abs() / peak).max().item())
            print("chunk_kb %d utt %d: log err strong bands ours-ref32 %.2e ours-fp64 %.2e ref32-fp64 %.2e | linear/peak ours %.2e ref %.2e"
                  % (ckb, b, e["o32"], e["o64"], e["r64"], lin["o64"], lin["r64"]))
            for k in worst:
                worst[k] = max(worst[k], e[k])
        print("chunk_kb %d worst: %s" % (ckb, worst))


if __name__ == "__main__":
    main()
