"""CPU oracle for the log-mel front-end and the collate.  TEST INFRASTRUCTURE ONLY.

Restates ``nntts.datasets.meldataset.mel_spectrogram`` (datasets/meldataset.py:49-82) and
``nntts.datasets.taco2_data.TextMelCollate.__call__`` (datasets/taco2_data.py:95-139) with the same torch CPU ops in
the same order.  Only ``tests/`` and ``bench.py``'s CPU legs may import it.

Where the arithmetic lives: ``torch.stft`` (this image's torch 2.11 CPU FFT) and ``librosa.filters.mel``.  librosa is
NOT installed in this image and is not vendored in /root/reference (``setup.py`` lists ``librosa`` unpinned), so the
mel filter bank is restated here from librosa's published algorithm for the arguments the reference passes --
``mel(sr, n_fft, n_mels, fmin, fmax)`` with the defaults ``htk=False`` (Slaney mel scale: linear below 1 kHz, log
above) and ``norm='slaney'`` (area normalisation ``2 / (f[i+2] - f[i])``), computed in float64 and returned as
float32 like librosa does.  Parity pinning: ``tests/golden/make_golden_frontend.py`` imports the UNMODIFIED reference
module with that function injected as ``librosa.filters.mel`` (and a shim that passes ``return_complex`` to
``torch.stft``, which the 2020 call site predates) and commits its outputs as ``tests/golden/frontend_*.npz``; the
filter bank itself is additionally pinned by closed-form checks (``tests/test_oracle_golden.py``).  It is therefore
pinned to the reference's own code path, not to a real librosa install: "filter-bank parity unpinned against librosa".
"""
from __future__ import annotations

import numpy as np
import torch

N_FFT, NUM_MELS, SAMPLING_RATE, HOP_SIZE, WIN_SIZE, FMIN, FMAX = 1024, 80, 22050, 256, 1024, 0, 8000


# --------------------------------------------------------------------------- librosa.filters.mel (restated)
def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore"):
        log_t = f >= min_log_hz
        mels = np.where(log_t, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)
    return mels


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = m >= min_log_mel
    return np.where(log_t, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel_basis(sr=SAMPLING_RATE, n_fft=N_FFT, n_mels=NUM_MELS, fmin=FMIN, fmax=FMAX):
    """``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` with htk=False, norm='slaney' -> float32 [n_mels, 1 + n_fft//2]."""
    if fmax is None:
        fmax = sr / 2.0
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2, dtype=np.float64)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


# --------------------------------------------------------------------------- mel_spectrogram
def mel_spectrogram(y, n_fft=N_FFT, num_mels=NUM_MELS, sampling_rate=SAMPLING_RATE, hop_size=HOP_SIZE,
                    win_size=WIN_SIZE, fmin=FMIN, fmax=FMAX, center=False):
    """datasets/meldataset.py:49-82: y float32 [B, L] in [-1, 1] -> log-mel float32 [B, num_mels, frames]."""
    mel = torch.from_numpy(slaney_mel_basis(sampling_rate, n_fft, num_mels, fmin, fmax)).float()
    window = torch.hann_window(win_size)
    pad = int((n_fft - hop_size) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)          # :66-67
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=window, center=center,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)  # :69-70
    spec = torch.view_as_real(spec)                   # the [..., 2] real view the 2020 call site received
    spec = torch.sqrt(spec.pow(2).sum(-1) + (1e-9))   # :72
    spec = torch.matmul(mel, spec)                    # :74
    return torch.log(torch.clamp(spec, min=1e-5) * 1)  # :75 -> dynamic_range_compression_torch :29-30


def num_frames(length, n_fft=N_FFT, hop_size=HOP_SIZE):
    """Frames of an utterance of ``length`` samples: reflect padding of (n_fft - hop) / 2 on both sides, center=False."""
    return 1 + (length + 2 * int((n_fft - hop_size) / 2) - n_fft) // hop_size


# --------------------------------------------------------------------------- TextMelCollate
def text_mel_collate(batch, n_frames_per_step=1):
    """datasets/taco2_data.py:101-139: batch = list of (text int tensor [L1], mel float [num_mels, T]) ->
    (text_padded int64 [B, L1max], input_lengths int64 [B] (sorted, descending), mel_padded float32 [B, Tmax, num_mels],
    output_lengths int64 [B])."""
    input_lengths, ids = torch.sort(torch.LongTensor([len(x[0]) for x in batch]), dim=0, descending=True)
    max_input_len = int(input_lengths[0])
    text_padded = torch.zeros(len(batch), max_input_len, dtype=torch.int64)
    for i in range(len(ids)):
        text = batch[ids[i]][0]
        text_padded[i, :text.size(0)] = text
    num_mels = batch[0][1].size(0)
    max_target_len = max(x[1].size(1) for x in batch)
    if max_target_len % n_frames_per_step != 0:
        max_target_len += n_frames_per_step - max_target_len % n_frames_per_step
    mel_padded = torch.zeros(len(batch), num_mels, max_target_len, dtype=torch.float32)
    output_lengths = torch.zeros(len(batch), dtype=torch.int64)
    for i in range(len(ids)):
        mel = batch[ids[i]][1]
        mel_padded[i, :, :mel.size(1)] = mel
        output_lengths[i] = mel.size(1)
    return text_padded, input_lengths, mel_padded.transpose(1, 2), output_lengths


# --------------------------------------------------------------------------- synthetic audio
def make_audio(seed, lengths):
    """Speech-like synthetic waveforms in [-1, 1]: a few harmonics with a slow envelope plus low-level noise, so that
    the spectrum has the dynamic range of real recordings (loud low bins, quiet high bins).  float32 [B, max(lengths)],
    zero beyond each length."""
    g = torch.Generator().manual_seed(int(seed))
    B, L = len(lengths), int(max(lengths))
    t = torch.arange(L, dtype=torch.float64) / SAMPLING_RATE
    y = torch.zeros(B, L, dtype=torch.float64)
    for b in range(B):
        f0 = 90.0 + 160.0 * torch.rand(1, generator=g).item()
        for h in range(1, 9):
            amp = 0.25 / h * (0.5 + torch.rand(1, generator=g).item())
            ph = 6.283185307179586 * torch.rand(1, generator=g).item()
            y[b] += amp * torch.sin(6.283185307179586 * f0 * h * t + ph)
        env = 0.55 + 0.45 * torch.sin(6.283185307179586 * (1.5 + torch.rand(1, generator=g).item()) * t)
        y[b] = y[b] * env + 0.003 * torch.randn(L, generator=g, dtype=torch.float64)
        y[b, int(lengths[b]):] = 0.0
    return y.clamp(-1.0, 1.0).float()
