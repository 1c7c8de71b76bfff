#!/usr/bin/env python3
"""Golden fixtures for the log-mel front-end and the collate from the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``):  ``python tests/golden/make_golden_frontend.py``

``nntts/datasets/meldataset.py`` is imported as the file it is.  Two things it needs do not exist in this image:
  * ``librosa`` (``from librosa.util import normalize``, ``from librosa.filters import mel``): a stand-in module is
    registered whose ``filters.mel`` is ``oracle.frontend_oracle.slaney_mel_basis`` -- librosa's published algorithm for
    the defaults the reference relies on (htk=False, norm='slaney'), restated; ``util.normalize`` is never called on
    the path;
  * ``torch.stft`` without ``return_complex`` (the 2020 call site, datasets/meldataset.py:69): torch 2.11 refuses real
    inputs without it, so for the duration of the call ``torch.stft`` is wrapped to pass ``return_complex=True`` and
    hand back the ``[..., 2]`` real view the old API returned.
``TextMelCollate`` (datasets/taco2_data.py:95-139) is executed from its own source text, cut out of the file, because
importing ``taco2_data`` pulls in ``nntts.text`` -> ``unidecode`` (absent); the class body has no other dependency.
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
warnings.filterwarnings("ignore")

from oracle import frontend_oracle as fo  # noqa: E402

REF = "/root/reference/nntts/datasets"


def load_reference_meldataset():
    lib = types.ModuleType("librosa")
    lib.util = types.ModuleType("librosa.util")
    lib.util.normalize = lambda x, *a, **k: x
    lib.filters = types.ModuleType("librosa.filters")
    lib.filters.mel = lambda sr, n_fft, n_mels, fmin, fmax: fo.slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)
    sys.modules.setdefault("librosa", lib)
    sys.modules.setdefault("librosa.util", lib.util)
    sys.modules.setdefault("librosa.filters", lib.filters)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_meldataset", os.path.join(REF, "meldataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_mel(mod, y):
    orig = torch.stft

    def stft_2020(*a, **k):
        k.setdefault("return_complex", True)
        return torch.view_as_real(orig(*a, **k))
    torch.stft = stft_2020
    try:
        mod.mel_basis.clear()
        mod.hann_window.clear()
        return mod.mel_spectrogram(y)
    finally:
        torch.stft = orig


def load_reference_collate():
    src = open(os.path.join(REF, "taco2_data.py"), encoding="utf-8").read()
    body = src[src.index("class TextMelCollate"):]
    ns = {"torch": torch}
    exec(compile(body, os.path.join(REF, "taco2_data.py"), "exec"), ns)
    return ns["TextMelCollate"]


def main():
    mod = load_reference_meldataset()
    # 1. per-utterance mels the way TextMelLoader.get_mel calls it (datasets/taco2_data.py:72-78): [1, L] -> [80, T]
    lengths = [22050, 9000, 256 * 40, 256 * 5 + 17, 1300]
    audio = fo.make_audio(7, lengths)
    mels = []
    for b, L in enumerate(lengths):
        m = reference_mel(mod, audio[b:b + 1, :L]).squeeze(0)
        assert m.shape == (80, fo.num_frames(L)), (m.shape, fo.num_frames(L))
        mels.append(m)
    out = {"seed": np.int64(7), "lengths": np.array(lengths, dtype=np.int64)}
    for b, m in enumerate(mels):
        out["mel_%d" % b] = m.numpy()
    np.savez_compressed(os.path.join(HERE, "frontend_mel.npz"), **out)
    # 2. collate of (text, mel) pairs
    Collate = load_reference_collate()
    g = torch.Generator().manual_seed(11)
    t_lens = [13, 40, 7, 40, 22]
    texts = [torch.randint(1, 76, (n,), generator=g).long() for n in t_lens]
    tp, il, mp, ol = Collate()([(t, m) for t, m in zip(texts, mels)])
    np.savez_compressed(os.path.join(HERE, "frontend_collate.npz"), seed=np.int64(11),
                        t_lens=np.array(t_lens, dtype=np.int64), text_padded=tp.numpy(), input_lengths=il.numpy(),
                        mel_padded=mp.contiguous().numpy(), output_lengths=ol.numpy(),
                        **{"text_%d" % i: t.numpy() for i, t in enumerate(texts)})
    # the oracle restatement agrees with what was just produced
    for b, L in enumerate(lengths):
        o = fo.mel_spectrogram(audio[b:b + 1, :L]).squeeze(0)
        assert torch.equal(o, mels[b]), (b, (o - mels[b]).abs().max())
    r = fo.text_mel_collate([(t, m) for t, m in zip(texts, mels)])
    assert all(torch.equal(a, b) for a, b in zip(r, (tp, il, mp, ol)))
    print("wrote frontend_mel.npz, frontend_collate.npz;", [tuple(m.shape) for m in mels],
          "mel range %.2f .. %.2f" % (min(float(m.min()) for m in mels), max(float(m.max()) for m in mels)))


if __name__ == "__main__":
    main()
