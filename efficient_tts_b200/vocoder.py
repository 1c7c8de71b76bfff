"""Drop-in for ``nntts.vocoders.hifigan_model.Generator`` (``/root/reference/nntts/vocoders/hifigan_model.py``),
the HiFi-GAN V1 generator the reference runs right after ``model.inference`` (bin/inference.py:85,108-109):

    voc_model = load_hifigan_generator(device)          # Generator(h); load_state_dict; eval; remove_weight_norm
    y = voc_model(mel_pred.transpose(1, 2))             # [1, 80, T] -> [1, 1, T * 256]

Same constructor (``Generator(h)`` with the attribute-style config of vocoders/HiFiGAN_LJ_V1/config.json), same
parameter names (``conv_pre``, ``ups.i``, ``resblocks.n.convs1.m`` / ``convs2.m``, ``conv_post``; weight-normed until
``remove_weight_norm()``), same call.  The forward runs in ``libefts_b200.so`` (``efts_vocoder_*``, include/efts_b200.h);
there is no CPU path.
"""
import ctypes

import torch
from torch.nn.utils import remove_weight_norm, weight_norm

from . import _lib
from .engine import RANGE_MESSAGE, _ptr, fold_state_dict

LRELU_SLOPE = 0.1


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


def _cfg(h, name):
    return h[name] if isinstance(h, dict) else getattr(h, name)


class ResBlock1(torch.nn.Module):
    """Parameter holder with the reference's layout (vocoders/hifigan_model.py:31-54)."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.h = h
        self.convs1 = torch.nn.ModuleList([
            weight_norm(torch.nn.Conv1d(channels, channels, kernel_size, 1, dilation=d,
                                        padding=get_padding(kernel_size, d))) for d in dilation])
        self.convs2 = torch.nn.ModuleList([
            weight_norm(torch.nn.Conv1d(channels, channels, kernel_size, 1, dilation=1,
                                        padding=get_padding(kernel_size, 1))) for _ in dilation])
        for m in list(self.convs1) + list(self.convs2):
            m.weight_v.data.normal_(0.0, 0.01)          # init_weights (vocoders/utils.py)

    def remove_weight_norm(self):
        for m in list(self.convs1) + list(self.convs2):
            remove_weight_norm(m)


class ResBlock2(torch.nn.Module):
    """Parameter holder with the reference's layout (vocoders/hifigan_model.py:71-81)."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3)):
        super().__init__()
        self.h = h
        self.convs = torch.nn.ModuleList([
            weight_norm(torch.nn.Conv1d(channels, channels, kernel_size, 1, dilation=d,
                                        padding=get_padding(kernel_size, d))) for d in dilation])
        for m in self.convs:
            m.weight_v.data.normal_(0.0, 0.01)          # init_weights (vocoders/utils.py)

    def remove_weight_norm(self):
        for m in self.convs:
            remove_weight_norm(m)


class VocoderEngine:
    """One prepacked generator on one CUDA device."""

    def __init__(self, device, state_dict, h, options=None):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("efts_b200 runs on CUDA (sm_100a) devices only; got %s" % (self.device,))
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        rates, ksz = list(_cfg(h, "upsample_rates")), list(_cfg(h, "upsample_kernel_sizes"))
        rks, rds = list(_cfg(h, "resblock_kernel_sizes")), [list(d) for d in _cfg(h, "resblock_dilation_sizes")]
        rtype = 1 if str(_cfg(h, "resblock")) == "1" else 2           # vocoders/hifigan_model.py:102
        nd = len(rds[0])
        if len(rates) > 8 or len(rks) > 4 or nd > 3 or any(len(d) != nd for d in rds):
            raise NotImplementedError("generator topology outside the B200 path")
        self.hop = 1
        for u in rates:
            self.hop *= int(u)
        self.num_mels = 80                      # Conv1d(80, ...) is hard-coded at vocoders/hifigan_model.py:101
        cfg = _lib.EftsVocoderConfig()
        cfg.num_mels = self.num_mels
        cfg.upsample_initial_channel = int(_cfg(h, "upsample_initial_channel"))
        cfg.num_upsamples = len(rates)
        for i, (u, k) in enumerate(zip(rates, ksz)):
            cfg.upsample_rates[i], cfg.upsample_kernel_sizes[i] = int(u), int(k)
        cfg.num_kernels = len(rks)
        for j, (k, d) in enumerate(zip(rks, rds)):
            cfg.resblock_kernel_sizes[j] = int(k)
            for m in range(nd):
                cfg.resblock_dilations[j][m] = int(d[m])
        cfg.resblock_type, cfg.num_dilations = rtype, nd
        cfg.device = idx
        h_ctx = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.efts_vocoder_create(ctypes.byref(cfg), ctypes.byref(h_ctx)))
            self._h = h_ctx
            for name, value in (options or {}).items():          # packing options are read by finalize
                _lib.check(self.lib.efts_set_option(self._h, name.encode(), int(value)))
            for name, t in fold_state_dict(state_dict).items():
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                _lib.check(self.lib.efts_set_weight(self._h, name.encode(), _ptr(t), shape, t.dim()))
            _lib.check(self.lib.efts_vocoder_finalize(self._h))
        self._ws = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.efts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        _lib.check(self.lib.efts_set_option(self._h, name.encode(), int(value)))

    def launch_count(self):
        return int(self.lib.efts_launch_count(self._h))

    def forward(self, mel, check_numerics=True):
        if mel.device != self.device:
            raise RuntimeError("mel is on %s but the generator is on %s" % (mel.device, self.device))
        if mel.dim() != 3 or mel.shape[1] != self.num_mels:
            raise RuntimeError("expected input [B, %d, T] (vocoders/hifigan_model.py:101), got %s" % (
                self.num_mels, tuple(mel.shape)))
        x = mel.to(torch.float32).contiguous()
        B, _, T = x.shape
        with torch.cuda.device(self.device):
            n = int(self.lib.efts_vocoder_workspace_bytes(self._h, B, T))
            if self._ws is None or self._ws.numel() < n:
                self._ws = None
                self._ws = torch.empty(n, dtype=torch.uint8, device=self.device)
            y = torch.empty(B, 1, T * self.hop, dtype=torch.float32, device=self.device)
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.efts_vocoder_forward(self._h, _ptr(x), B, T, _ptr(y), _ptr(self._ws),
                                                     self._ws.numel(), st))
            if check_numerics:
                flags = ctypes.c_int32(0)
                _lib.check(self.lib.efts_error_flags(self._h, st, ctypes.byref(flags)))
                if flags.value & 8:
                    raise FloatingPointError(RANGE_MESSAGE + " [flags 0x%x]" % flags.value)
        return y


class Generator(torch.nn.Module):
    """HiFi-GAN generator, forward on B200.  Constructor and parameter layout: vocoders/hifigan_model.py:95-118."""

    def __init__(self, h):
        super().__init__()
        self.h = h
        rks, rds = _cfg(h, "resblock_kernel_sizes"), _cfg(h, "resblock_dilation_sizes")
        rates, ksz = _cfg(h, "upsample_rates"), _cfg(h, "upsample_kernel_sizes")
        c0 = _cfg(h, "upsample_initial_channel")
        resblock = ResBlock1 if str(_cfg(h, "resblock")) == "1" else ResBlock2      # :102
        self.num_kernels = len(rks)
        self.num_upsamples = len(rates)
        self.conv_pre = weight_norm(torch.nn.Conv1d(80, c0, 7, 1, padding=3))
        self.ups = torch.nn.ModuleList()
        for i, (u, k) in enumerate(zip(rates, ksz)):
            self.ups.append(weight_norm(torch.nn.ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u,
                                                                padding=(k - u) // 2)))
        self.resblocks = torch.nn.ModuleList()
        ch = c0
        for i in range(len(self.ups)):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(rks, rds):
                self.resblocks.append(resblock(h, ch, k, d))
        self.conv_post = weight_norm(torch.nn.Conv1d(ch, 1, 7, 1, padding=3))
        for m in list(self.ups) + [self.conv_post]:
            m.weight_v.data.normal_(0.0, 0.01)          # init_weights

    # ------------------------------------------------------------------ engine plumbing
    def _get_engine(self):
        from .layers import _fingerprint
        fp = _fingerprint(self)
        eng = self.__dict__.get("_efts_engine")
        if eng is None or self.__dict__.get("_efts_fp") != fp:
            if eng is not None:
                eng.close()
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("efficient_tts_b200 modules compute on a CUDA sm_100a device only (parameters are "
                                   "on %s); there is no CPU path -- call .to('cuda')" % dev)
            eng = VocoderEngine(dev, self.state_dict(), self.h, self.__dict__.get("_efts_options"))
            self.__dict__["_efts_engine"] = eng
            self.__dict__["_efts_fp"] = fp
        return eng

    def set_engine_options(self, **options):
        """Packing / launch options of the C library (``efts_set_option``), applied when the engine is (re)built."""
        self.__dict__["_efts_options"] = dict(options)
        eng = self.__dict__.pop("_efts_engine", None)
        if eng is not None:
            eng.close()

    # ------------------------------------------------------------------ reference surface
    def forward(self, x):
        """vocoders/hifigan_model.py:120-136: mel [B, 80, T] -> waveform [B, 1, T * prod(upsample_rates)]."""
        if self.training:
            raise RuntimeError("efficient_tts_b200 is a forward-only engine: call .eval() first")
        return self._get_engine().forward(x)

    def remove_weight_norm(self):
        """vocoders/hifigan_model.py:138-145."""
        for m in self.ups:
            remove_weight_norm(m)
        for m in self.resblocks:
            m.remove_weight_norm()
        remove_weight_norm(self.conv_pre)
        remove_weight_norm(self.conv_post)


def load_hifigan_generator(device, config_path, ckpt_path):
    """``nntts.vocoders.hifigan_model.load_hifigan_generator`` (vocoders/hifigan_model.py:18-28) with explicit paths
    (the reference reads ``HiFiGAN_LJ_V1/config.json`` and ``generator_v1`` next to its own module): build the
    generator from the JSON config, load ``state_dict["generator"]``, ``eval()``, ``remove_weight_norm()``."""
    import json
    with open(config_path) as f:
        h = json.loads(f.read())

    class _AttrDict(dict):
        __getattr__ = dict.__getitem__

    generator = Generator(_AttrDict(h))
    state = torch.load(ckpt_path, map_location="cpu")
    generator.load_state_dict(state["generator"])
    generator.eval()
    generator.remove_weight_norm()
    return generator.to(device)
