"""Log-mel front-end and collate on the GPU.

Host-side mirror of ``nntts.datasets.meldataset.mel_spectrogram`` (datasets/meldataset.py:49-82) and
``nntts.datasets.taco2_data.TextMelCollate`` (datasets/taco2_data.py:95-139): same names, arguments and return
layouts; the arithmetic runs in ``libefts_b200.so`` (``efts_frontend_*``, include/efts_b200.h).  There is no CPU path.

The reference's ``mel_spectrogram`` takes its filter bank from ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)``
(defaults htk=False, norm='slaney').  librosa is not a dependency of this package; ``slaney_mel_basis`` below follows
librosa's published algorithm for exactly that call.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from .engine import RANGE_MESSAGE, _ptr


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    logstep = math.log(6.4) / 27.0
    lin = f / f_sp
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    logstep = math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax):
    """float32 [n_mels, n_fft // 2 + 1]: triangular filters on the Slaney mel scale, area-normalised."""
    fmax = sr / 2.0 if fmax is None else fmax
    freqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - freqs[None, :]
    w = np.maximum(0.0, np.minimum(-ramps[:-2] / width[:-1, None], ramps[2:] / width[1:, None]))
    w *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return w.astype(np.float32)


def windowed_dft_basis(n_fft, hop):
    """fp32 [n_fft, hop, n_fft // hop] in the GEMM's [N, K, taps] layout: output column n of frame t is
    sum_{tap, k} chunk[t + tap, k] * W[n, k, tap].  Columns 0 .. n_fft/2: w[s] cos(2 pi n s / n_fft) (real parts);
    columns n_fft/2 + 1 .. n_fft - 1: -w[s] sin(2 pi (n - n_fft/2) s / n_fft) (imaginary parts of bins 1 .. n_fft/2 - 1)
    with s = tap * hop + k and w the periodic Hann window ``torch.hann_window(n_fft)`` (:65)."""
    half = n_fft // 2
    s = np.arange(n_fft, dtype=np.float64)
    window = torch.hann_window(n_fft, dtype=torch.float64).numpy()
    n_re = np.arange(half + 1, dtype=np.float64)
    n_im = np.arange(1, half, dtype=np.float64)
    # exact argument reduction: (n * s) mod n_fft keeps cos / sin arguments in [0, 2 pi)
    re = np.cos(2.0 * np.pi * np.mod(np.outer(n_re, s), n_fft) / n_fft) * window[None, :]
    im = -np.sin(2.0 * np.pi * np.mod(np.outer(n_im, s), n_fft) / n_fft) * window[None, :]
    w = np.concatenate([re, im], axis=0)                      # [n_fft, n_fft] = [N, s]
    return np.ascontiguousarray(w.reshape(n_fft, n_fft // hop, hop).transpose(0, 2, 1)).astype(np.float32)


class LogMelFrontend:
    """One prepacked front-end (windowed DFT basis + mel filter bank) on one CUDA device."""

    def __init__(self, device, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0,
                 fmax=8000):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the log-mel front-end runs on CUDA (sm_100a) devices only; got %s" % (self.device,))
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.n_fft, self.hop, self.num_mels = int(n_fft), int(hop_size), int(num_mels)
        cfg = _lib.EftsFrontendConfig(int(n_fft), int(hop_size), int(win_size), int(num_mels), idx)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.efts_frontend_create(ctypes.byref(cfg), ctypes.byref(h)))
            self._h = h
            kp = (n_fft // 2 + 1 + 7) // 8 * 8
            mel = np.zeros((num_mels, kp), dtype=np.float32)
            mel[:, : n_fft // 2 + 1] = slaney_mel_basis(sampling_rate, n_fft, num_mels, fmin, fmax)
            weights = {"stft.weight": windowed_dft_basis(n_fft, hop_size), "stft.bias": np.zeros(n_fft, np.float32),
                       "mel_basis.weight": mel, "mel_basis.bias": np.zeros(num_mels, np.float32)}
            for name, arr in weights.items():
                t = torch.from_numpy(np.ascontiguousarray(arr))
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                _lib.check(self.lib.efts_set_weight(self._h, name.encode(), _ptr(t), shape, t.dim()))
            _lib.check(self.lib.efts_frontend_finalize(self._h))
        self._ws = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.efts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        _lib.check(self.lib.efts_set_option(self._h, name.encode(), int(value)))

    def frames(self, length):
        return int(self.lib.efts_frontend_frames(self._h, int(length)))

    def launch_count(self):
        return int(self.lib.efts_launch_count(self._h))

    def __call__(self, audio, lengths=None, check=True):
        """audio float [B, Lmax] on this device, lengths int64 [B] or None -> (mel fp32 [B, Tmax, num_mels],
        mel_lengths int64 [B]); mel is zero beyond each utterance's frames."""
        if audio.device != self.device:
            raise RuntimeError("audio is on %s but the front-end is on %s" % (audio.device, self.device))
        y = audio.to(torch.float32).contiguous()
        if y.dim() != 2:
            raise RuntimeError("audio must be [B, L]")
        B, L = y.shape
        tmax = self.frames(L)
        if tmax < 1:
            raise RuntimeError("Argument #4: Padding size should be less than the corresponding input dimension "
                               "(%d samples are too short for the %d-sample reflect padding)" % (L, (self.n_fft - self.hop) // 2))
        lens = None if lengths is None else lengths.to(self.device, torch.int64).contiguous()
        with torch.cuda.device(self.device):
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            mel = torch.empty(B, tmax, self.num_mels, dtype=torch.float32, device=self.device)
            mel_lens = torch.empty(B, dtype=torch.int64, device=self.device)
            n = int(self.lib.efts_frontend_workspace_bytes(self._h, B, L))
            if self._ws is None or self._ws.numel() < n:
                self._ws = None
                self._ws = torch.empty(n, dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.efts_frontend_forward(self._h, _ptr(y), _ptr(lens), B, L, _ptr(mel), _ptr(mel_lens),
                                                      _ptr(self._ws), self._ws.numel(), st))
            if check:
                flags = ctypes.c_int32(0)
                _lib.check(self.lib.efts_error_flags(self._h, st, ctypes.byref(flags)))
                if flags.value & 16:
                    raise RuntimeError("a length lies outside [0, %d]" % L)
                if flags.value & 32:
                    raise RuntimeError("Argument #4: Padding size should be less than the corresponding input dimension "
                                       "(an utterance is not longer than the reflect padding)")
                if flags.value & 8:
                    raise FloatingPointError(RANGE_MESSAGE)
        return mel, mel_lens


_FRONTENDS = {}


def _frontend_for(device, **kw):
    key = (str(device),) + tuple(sorted(kw.items()))
    fe = _FRONTENDS.get(key)
    if fe is None:
        fe = _FRONTENDS[key] = LogMelFrontend(device, **kw)
    return fe


def mel_spectrogram(y, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0, fmax=8000,
                    center=False):
    """datasets/meldataset.py:49-82 with the reference's signature: ``y`` float [B, L] in [-1, 1] on a CUDA device ->
    log-mel float32 [B, num_mels, frames] (a transposed view of the channels-last tensor the kernels write)."""
    if center:
        raise NotImplementedError("center=True is outside the path (the reference calls it with center=False)")
    if torch.min(y) < -1.:
        print('min value is ', torch.min(y))
    if torch.max(y) > 1.:
        print('max value is ', torch.max(y))
    fe = _frontend_for(y.device, n_fft=n_fft, num_mels=num_mels, sampling_rate=sampling_rate, hop_size=hop_size,
                       win_size=win_size, fmin=fmin, fmax=fmax)
    mel, _ = fe(y)
    return mel.transpose(1, 2)


class TextMelCollate:
    """datasets/taco2_data.py:95-139 for pairs whose mel already exists: ``batch`` = list of ``(text [L1], mel
    [num_mels, T])`` -> ``(text_padded int64 [B, L1max], input_lengths int64 [B] sorted descending, mel_padded fp32
    [B, Tmax, num_mels], output_lengths int64 [B])``.  Pure data movement (the reference runs it in DataLoader
    workers); tensors stay on the device their inputs are on."""

    def __init__(self, n_frames_per_step=1):
        self.n_frames_per_step = n_frames_per_step

    def __call__(self, batch):
        dev = batch[0][1].device
        input_lengths, order = torch.sort(torch.LongTensor([len(x[0]) for x in batch]), dim=0, descending=True)
        text_padded = torch.zeros(len(batch), int(input_lengths[0]), dtype=torch.int64, device=batch[0][0].device)
        num_mels = batch[0][1].size(0)
        tmax = max(x[1].size(1) for x in batch)
        if tmax % self.n_frames_per_step != 0:
            tmax += self.n_frames_per_step - tmax % self.n_frames_per_step
        mel_padded = torch.zeros(len(batch), num_mels, tmax, dtype=torch.float32, device=dev)
        output_lengths = torch.zeros(len(batch), dtype=torch.int64)
        for i, j in enumerate(order.tolist()):
            text, mel = batch[j]
            text_padded[i, : text.size(0)] = text
            mel_padded[i, :, : mel.size(1)] = mel
            output_lengths[i] = mel.size(1)
        return text_padded, input_lengths.to(text_padded.device), mel_padded.transpose(1, 2), output_lengths.to(dev)


class AudioTextCollate:
    """The loader + collate of the reference fused on the GPU: ``batch`` = list of ``(text int [L1], audio float [L])``
    (what ``TextMelLoader`` holds before ``get_mel``, datasets/taco2_data.py:46-78) -> exactly the tuple
    ``TextMelCollate`` returns for the same utterances, on ``device``: sorted by text length (descending), text
    zero-padded, and the log-mel of every utterance computed in ONE batched front-end call -- each utterance
    reflect-padded at its own length, frames beyond its own count zero."""

    def __init__(self, device, n_frames_per_step=1, **mel_kwargs):
        if n_frames_per_step != 1:
            raise NotImplementedError("n_frames_per_step != 1 is outside the path (the recipe uses 1)")
        self.device = torch.device(device)
        self.frontend = _frontend_for(self.device, **({"n_fft": 1024, "num_mels": 80, "sampling_rate": 22050,
                                                       "hop_size": 256, "win_size": 1024, "fmin": 0, "fmax": 8000} | mel_kwargs))

    def __call__(self, batch):
        input_lengths, order = torch.sort(torch.LongTensor([len(x[0]) for x in batch]), dim=0, descending=True)
        order = order.tolist()
        B = len(batch)
        text_padded = torch.zeros(B, int(input_lengths[0]), dtype=torch.int64)
        a_lens = torch.tensor([batch[j][1].numel() for j in order], dtype=torch.int64)
        audio = torch.zeros(B, int(a_lens.max()), dtype=torch.float32).pin_memory()
        for i, j in enumerate(order):
            text, wav = batch[j]
            text_padded[i, : text.numel()] = text
            audio[i, : wav.numel()] = wav
        mel, mel_lens = self.frontend(audio.to(self.device, non_blocking=True), a_lens.to(self.device))
        return text_padded.to(self.device), input_lengths.to(self.device), mel, mel_lens
