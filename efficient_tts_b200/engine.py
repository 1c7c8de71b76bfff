"""Thin Python owner of one ``efts_ctx``: weight marshalling, workspace caching and raw-pointer calls.

PyTorch is used for device memory and streams only; every computation below is a call into
``libefts_b200.so``.  Reference citations are relative to /root/reference/nntts.
"""
import ctypes

import torch

from . import _lib

N_CHANNELS = 512
STACK_TIMEOUT_MESSAGE = ("the resident layer-stack kernel gave up waiting at a grid barrier (its CTAs were not all "
                         "resident); the results of this call are invalid (include/efts_b200.h, flags bit 6)")
RANGE_MESSAGE = ("an activation left the fp16 operand range (|x| > 65504): the split-fp16 tensor-core "
                 "scheme cannot represent it (include/efts_b200.h, flags bit 3)")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def fold_state_dict(state_dict):
    """Reference ``state_dict`` -> folded fp32 CPU tensors keyed like ``remove_weight_norm()``
    leaves them (layers/efts_modules.py:81-99): ``*.weight_g`` / ``*.weight_v`` pairs become
    ``*.weight`` through torch's own ``_weight_norm`` so the effective weights are bit-identical
    to the ones the reference convolves with."""
    out = {}
    for k, v in state_dict.items():
        if k.endswith(".weight_g"):
            p = k[: -len(".weight_g")]
            g = v.detach().to("cpu", torch.float32)
            w = state_dict[p + ".weight_v"].detach().to("cpu", torch.float32)
            out[p + ".weight"] = torch._weight_norm(w, g, 0).contiguous()
        elif k.endswith(".weight_v"):
            continue
        elif k.endswith("parametrizations.weight.original0") or k.endswith("parametrizations.weight.original1"):
            raise NotImplementedError("parametrize-style weight norm checkpoints are not supported; "
                                      "use torch.nn.utils.weight_norm keys (weight_g / weight_v)")
        else:
            out[k] = v.detach().to("cpu", torch.float32).contiguous()
    return out


class Engine:
    """One prepacked model on one CUDA device."""

    def __init__(self, device, state_dict, *, num_symbols, odim=80, n_channels=N_CHANNELS, k_size=5,
                 n_text_encoder_layer=5, n_mel_encoder_layer=3, n_decoder_layer=6, n_duration_layer=2,
                 duration_kernel_size=3, sigma=0.01, sigma_e=0.5, duration_offset=1.0,
                 leaky_relu_slope=0.1, use_masking=True):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("efts_b200 runs on CUDA (sm_100a) devices only; got %s" % (self.device,))
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.odim, self.C = int(odim), int(n_channels)
        cfg = _lib.EftsConfig(int(num_symbols), int(odim), int(n_channels), int(k_size),
                              int(n_text_encoder_layer), int(n_mel_encoder_layer), int(n_decoder_layer),
                              int(n_duration_layer), int(duration_kernel_size), float(sigma),
                              float(sigma_e), float(duration_offset), float(leaky_relu_slope),
                              int(bool(use_masking)), idx)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.efts_create(ctypes.byref(cfg), ctypes.byref(h)))
            self._h = h
            folded = fold_state_dict(state_dict)
            for name, t in folded.items():
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                _lib.check(self.lib.efts_set_weight(self._h, name.encode(), _ptr(t), shape, t.dim()))
            _lib.check(self.lib.efts_finalize_weights(self._h))
        self._ws = None
        self._t2_dev = None
        self._host_words = (ctypes.c_int32 * 8)()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.efts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def set_option(self, name, value):
        _lib.check(self.lib.efts_set_option(self._h, name.encode(), int(value)))

    def launch_count(self):
        return int(self.lib.efts_launch_count(self._h))

    def profile_enable(self, tag_mask):
        """Bracket launches of the tagged kinds with CUDA events (include/efts_b200.h)."""
        _lib.check(self.lib.efts_profile_enable(self._h, int(tag_mask)))

    def profile_read(self, tag):
        """(total milliseconds, launches) recorded for one tag since ``profile_enable``."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        _lib.check(self.lib.efts_profile_read(self._h, int(tag), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def profile_kernel_name(self, tag):
        """Instantiation of the last tensor-core GEMM launched under ``tag`` while profiling was enabled."""
        buf = ctypes.create_string_buffer(96)
        _lib.check(self.lib.efts_profile_kernel_name(self._h, int(tag), buf, 96))
        return buf.value.decode()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def workspace_for(self, B, T1, T2):
        n = int(self.lib.efts_workspace_bytes(self._h, int(B), int(T1), int(T2)))
        return self.workspace(n), n

    def _f32(self, t, name):
        if t.device != self.device:
            raise RuntimeError("%s is on %s but the model is on %s" % (name, t.device, self.device))
        return t.to(torch.float32).contiguous()

    def _i64(self, t, name):
        if t.device != self.device:
            raise RuntimeError("%s is on %s but the model is on %s" % (name, t.device, self.device))
        return t.to(torch.int64).contiguous()

    # ------------------------------------------------------------------ model-level calls
    def forward(self, text, text_lengths, speech, speech_lengths):
        """models/efficient_tts.py:120-228.  Returns (imv, reconst_alpha, mel_pred, scalars[8])."""
        text = self._i64(text, "text")
        tl = self._i64(text_lengths, "text_lengths")
        speech = self._f32(speech, "speech")
        sl = self._i64(speech_lengths, "speech_lengths")
        B, T1 = text.shape
        T2 = speech.shape[1]
        if speech.shape[0] != B or speech.shape[2] != self.odim or tl.numel() != B or sl.numel() != B:
            raise RuntimeError("inconsistent batch shapes: text %s speech %s lengths %s %s" % (
                tuple(text.shape), tuple(speech.shape), tuple(tl.shape), tuple(sl.shape)))
        with torch.cuda.device(self.device):
            imv = torch.empty(B, T2, dtype=torch.float32, device=self.device)
            ra = torch.empty(B, T1, T2, dtype=torch.float32, device=self.device)
            mel = torch.empty(B, T2, self.odim, dtype=torch.float32, device=self.device)
            scal = torch.empty(8, dtype=torch.float32, device=self.device)
            ws, n = self.workspace_for(B, T1, T2)
            _lib.check(self.lib.efts_forward(self._h, _ptr(text), _ptr(tl), _ptr(speech), _ptr(sl), B, T1, T2,
                                             _ptr(imv), _ptr(ra), _ptr(mel), _ptr(scal), _ptr(ws), n,
                                             self._stream()))
        return imv, ra, mel, scal

    def inference(self, text, check_numerics=True):
        """models/efficient_tts.py:230-285, B = 1.  Returns (mel_pred[1,T2,odim], reconst_alpha[1,T1,T2])."""
        text = self._i64(text, "text")
        if text.dim() != 2 or text.shape[0] != 1:
            # the reference's `.item()` at models/efficient_tts.py:361 only works for one utterance
            raise RuntimeError("a Tensor with %d elements cannot be converted to Scalar (inference is "
                               "B=1 like the reference, models/efficient_tts.py:361)" % text.shape[0])
        T1 = text.shape[1]
        with torch.cuda.device(self.device):
            if self._t2_dev is None:
                self._t2_dev = torch.empty(2, dtype=torch.int32, device=self.device)
            t2_dev = self._t2_dev
            # phase 1 only touches the text-sized part of the workspace
            ws, n = self.workspace_for(1, T1, 1)
            _lib.check(self.lib.efts_inference_phase1(self._h, _ptr(text), T1, _ptr(t2_dev), _ptr(ws), n,
                                                      self._stream()))
            # the reference's .item() sync (:361), through the context's pinned staging words
            _lib.check(self.lib.efts_read_words(self._h, _ptr(t2_dev), self._host_words, 2, self._stream()))
            t2, flags = self._host_words[0], self._host_words[1]
            if flags & 4:
                raise IndexError("index out of range in self")   # embedding lookup, :246
            if flags & 8:
                raise FloatingPointError(RANGE_MESSAGE)
            if flags & 64:
                raise RuntimeError(STACK_TIMEOUT_MESSAGE)
            if t2 < 1:
                raise RuntimeError("predicted length T2=%d; the reference's decoder conv rejects an "
                                   "empty sequence" % t2)
            need = int(self.lib.efts_workspace_bytes(self._h, 1, T1, t2))
            if need > ws.numel():
                # grow, keeping phase 1's results (they live at the front of the buffer)
                big = torch.empty(need, dtype=torch.uint8, device=self.device)
                big[: ws.numel()].copy_(ws)
                self._ws = ws = big
            mel = torch.empty(1, t2, self.odim, dtype=torch.float32, device=self.device)
            ra = torch.empty(1, T1, t2, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.efts_inference_phase2(self._h, T1, t2, _ptr(mel), _ptr(ra), _ptr(ws),
                                                      ws.numel(), self._stream()))
            if check_numerics:
                self.check_error_flags()
        return mel, ra

    def inference_batch(self, text, text_lengths, check_numerics=True):
        """Ragged batched synthesis (include/efts_b200.h, efts_inference_batch_*): utterance b is computed
        exactly as ``inference(text[b:b+1, :text_lengths[b]])``.  Returns ``(mel_pred[B,T2max,odim],
        mel_lengths int64 [B], reconst_alpha[B,T1,T2max])``; frames beyond ``mel_lengths[b]`` are zero."""
        text = self._i64(text, "text")
        tl = self._i64(text_lengths, "text_lengths")
        if text.dim() != 2 or tl.numel() != text.shape[0]:
            raise RuntimeError("text must be [B, T1] and text_lengths [B]")
        B, T1 = text.shape
        tl_host = tl.cpu()
        if int(tl_host.min()) < 1 or int(tl_host.max()) > T1:
            raise RuntimeError("text_lengths must lie in [1, %d]" % T1)
        with torch.cuda.device(self.device):
            t2_dev = torch.empty(B + 1, dtype=torch.int32, device=self.device)
            ws, n = self.workspace_for(B, T1, 1)
            _lib.check(self.lib.efts_inference_batch_phase1(self._h, _ptr(text), _ptr(tl), B, T1, _ptr(t2_dev), _ptr(ws),
                                                            ws.numel(), self._stream()))
            host = t2_dev.cpu()                            # one read-back for the whole batch
            flags = int(host[B])
            if flags & 4:
                raise IndexError("index out of range in self")
            if flags & 8:
                raise FloatingPointError(RANGE_MESSAGE)
            t2 = host[:B].to(torch.int64)
            if int(t2.min()) < 1:
                raise RuntimeError("utterance %d has predicted length %d; the reference's decoder conv rejects an "
                                   "empty sequence" % (int(t2.argmin()), int(t2.min())))
            t2max = int(t2.max())
            need = int(self.lib.efts_workspace_bytes(self._h, B, T1, t2max))
            if need > ws.numel():
                big = torch.empty(need, dtype=torch.uint8, device=self.device)
                big[: ws.numel()].copy_(ws)
                self._ws = ws = big
            mel = torch.empty(B, t2max, self.odim, dtype=torch.float32, device=self.device)
            ra = torch.empty(B, T1, t2max, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.efts_inference_batch_phase2(self._h, B, T1, t2max, _ptr(t2_dev), _ptr(mel), _ptr(ra),
                                                            _ptr(ws), ws.numel(), self._stream()))
            # t2_dev must outlive the kernels that read it as the frame-length vector
            mel._efts_keepalive = t2_dev
            if check_numerics:
                self.check_error_flags()
        return mel, t2.to(self.device), ra

    def check_error_flags(self):
        """Synchronising read of the kernels' data-dependent error bits (include/efts_b200.h)."""
        flags = ctypes.c_int32(0)
        _lib.check(self.lib.efts_error_flags(self._h, self._stream(), ctypes.byref(flags)))
        if flags.value & 8:
            raise FloatingPointError(RANGE_MESSAGE)
        if flags.value & 64:
            raise RuntimeError(STACK_TIMEOUT_MESSAGE)
        return flags.value

    # ------------------------------------------------------------------ layer-level calls
    def conv_stack(self, stack, x_btc):
        x = self._f32(x_btc, "x")
        B, T, C = x.shape
        if C != self.C:
            raise RuntimeError("expected %d channels, got %d" % (self.C, C))
        with torch.cuda.device(self.device):
            y = torch.empty_like(x)
            ws, n = self.workspace_for(B, T, T)
            _lib.check(self.lib.efts_conv_stack_fwd(self._h, int(stack), _ptr(x), _ptr(y), B, T, _ptr(ws), n,
                                                    self._stream()))
        return y

    def duration_predictor(self, x_btc, lengths=None, mode=0):
        x = self._f32(x_btc, "xs")
        B, T, C = x.shape
        if C != self.C:
            raise RuntimeError("expected %d channels, got %d" % (self.C, C))
        with torch.cuda.device(self.device):
            out = torch.empty(B, T, dtype=torch.int64 if mode == 2 else torch.float32, device=self.device)
            lens = None if lengths is None else lengths.to(self.device, torch.int32).contiguous()
            ws, n = self.workspace_for(B, T, T)
            _lib.check(self.lib.efts_duration_predictor_fwd(self._h, _ptr(x), _ptr(lens), B, T, int(mode),
                                                            _ptr(out), _ptr(ws), n, self._stream()))
        return out

    def tap_gemm(self, x, w, ntaps=1, pad=0, batched=False):
        x = self._f32(x, "x")
        w = self._f32(w, "w")
        B, T, K = x.shape
        Z, N, K2 = w.shape
        assert K == K2 and Z == (B if batched else ntaps)
        with torch.cuda.device(self.device):
            out = torch.empty(B, T, N, dtype=torch.float32, device=self.device)
            n = (x.numel() + w.numel()) * 4 + 8192
            ws = self.workspace(n)
            _lib.check(self.lib.efts_tap_gemm(self._h, _ptr(x), _ptr(w), _ptr(out), B, T, K, N, int(ntaps),
                                              int(pad), int(bool(batched)), _ptr(ws), ws.numel(),
                                              self._stream()))
        return out

    def attention_alpha(self, query, key, text_lengths_i32):
        query, key = self._f32(query, "query"), self._f32(key, "key")
        B, T2, C = query.shape
        T1 = key.shape[1]
        with torch.cuda.device(self.device):
            alpha = torch.empty(B, T1, T2, dtype=torch.float32, device=self.device)
            ws, n = self.workspace_for(B, T1, T2)
            _lib.check(self.lib.efts_attention_alpha(self._h, _ptr(query), _ptr(key), _ptr(text_lengths_i32), B, T1, T2,
                                                     _ptr(alpha), _ptr(ws), n, self._stream()))
        return alpha

    def alignment(self, mel_h, key, value, text_lengths, speech_lengths):
        mel_h, key, value = self._f32(mel_h, "mel_h"), self._f32(key, "key"), self._f32(value, "value")
        B, T2, C = mel_h.shape
        T1 = key.shape[1]
        with torch.cuda.device(self.device):
            tl = text_lengths.to(self.device, torch.int32).contiguous()
            sl = speech_lengths.to(self.device, torch.int32).contiguous()
            imv = torch.empty(B, T2, dtype=torch.float32, device=self.device)
            e = torch.empty(B, T1, dtype=torch.float32, device=self.device)
            ra = torch.empty(B, T1, T2, dtype=torch.float32, device=self.device)
            ex = torch.empty(B, T2, C, dtype=torch.float32, device=self.device)
            ws, n = self.workspace_for(B, T1, T2)
            _lib.check(self.lib.efts_alignment_fwd(self._h, _ptr(mel_h), _ptr(key), _ptr(value), _ptr(tl),
                                                   _ptr(sl), B, T1, T2, _ptr(imv), _ptr(e), _ptr(ra), _ptr(ex),
                                                   _ptr(ws), n, self._stream()))
        return imv, e, ra, ex


def _stream_of(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def mask_lengths(mask):
    """bool prefix mask [B, T] (make_non_pad_mask output) -> int32 lengths [B], on the device."""
    lib = _lib.load()
    if mask.device.type != "cuda":
        raise RuntimeError("efts_b200 helpers run on CUDA tensors only")
    m = mask.to(torch.uint8).contiguous() if mask.dtype != torch.bool else mask.contiguous().view(torch.uint8)
    B, T = m.shape
    with torch.cuda.device(mask.device):
        out = torch.empty(B, dtype=torch.int32, device=mask.device)
        _lib.check(lib.efts_mask_lengths(_ptr(m), B, T, _ptr(out), _stream_of(mask.device)))
    return out


def index_vector(text_lengths_i32, T1):
    lib = _lib.load()
    dev = text_lengths_i32.device
    B = text_lengths_i32.numel()
    with torch.cuda.device(dev):
        p = torch.empty(B, T1, dtype=torch.float32, device=dev)
        _lib.check(lib.efts_index_vector(_ptr(text_lengths_i32), B, T1, _ptr(p), _stream_of(dev)))
    return p


def imv_generator(alpha, p, text_lengths_i32, speech_lengths_i32):
    lib = _lib.load()
    dev = alpha.device
    alpha = alpha.to(torch.float32).contiguous()
    p = p.to(torch.float32).contiguous()
    B, T1, T2 = alpha.shape
    with torch.cuda.device(dev):
        imv = torch.empty(B, T2, dtype=torch.float32, device=dev)
        ws = torch.empty(B * T2 * 4 + 1024, dtype=torch.uint8, device=dev)
        _lib.check(lib.efts_imv_generator(_ptr(alpha), _ptr(p), _ptr(text_lengths_i32), _ptr(speech_lengths_i32), B, T1,
                                          T2, _ptr(imv), _ptr(ws), ws.numel(), _stream_of(dev)))
    return imv


def aligned_positions(imv, p, text_lengths_i32, speech_lengths_i32, sigma_e):
    lib = _lib.load()
    dev = imv.device
    imv = imv.to(torch.float32).contiguous()
    p = p.to(torch.float32).contiguous()
    B, T2 = imv.shape
    T1 = p.shape[1]
    with torch.cuda.device(dev):
        e = torch.empty(B, T1, dtype=torch.float32, device=dev)
        _lib.check(lib.efts_aligned_positions(_ptr(imv), _ptr(p), _ptr(text_lengths_i32), _ptr(speech_lengths_i32), B, T1,
                                              T2, float(sigma_e), _ptr(e), _stream_of(dev)))
    return e


def reconstruct_alignment(e, delta, text_lengths_i32, speech_lengths_i32, T2):
    lib = _lib.load()
    dev = e.device
    e = e.to(torch.float32).contiguous()
    B, T1 = e.shape
    with torch.cuda.device(dev):
        out = torch.empty(B, T1, int(T2), dtype=torch.float32, device=dev)
        _lib.check(lib.efts_reconstruct_alignment(_ptr(e), _ptr(text_lengths_i32), _ptr(speech_lengths_i32), B, T1,
                                                  int(T2), float(delta), _ptr(out), _stream_of(dev)))
    return out


def length_regulator(xs, ds, ilens, alpha=1.0, pad_value=0.0, return_index=False):
    """layers/length_regulator.py:35-79 on the GPU (context-free entry points)."""
    lib = _lib.load()
    if xs.device.type != "cuda":
        raise RuntimeError("efts_b200 length regulator runs on CUDA tensors only")
    dev = xs.device
    x = xs.to(torch.float32).contiguous()
    B, T1 = ds.shape
    D = x[0, 0].numel()
    if ds.dtype != torch.int64 or not ds.is_contiguous() or ds.device != dev:
        if alpha == 1.0:
            raise RuntimeError("ds must be a contiguous int64 tensor on the same device as xs")
        ds = ds.to(dev, torch.int64).contiguous()
    il = ilens.to(dev, torch.int64).contiguous()
    with torch.cuda.device(dev):
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        ds_eff = torch.empty_like(ds)
        out_lens = torch.empty(B, dtype=torch.int64, device=dev)
        plan = torch.empty(2, dtype=torch.int64, device=dev)
        _lib.check(lib.efts_length_regulator_plan(_ptr(ds), _ptr(il), float(alpha), B, T1, _ptr(ds_eff),
                                                  _ptr(out_lens), _ptr(plan), st))
        tout, flags = (int(v) for v in plan.cpu())
        if flags & 1:
            raise RuntimeError("Trying to create tensor with negative dimension (negative duration)")
        out = torch.empty((B, tout) + tuple(x.shape[2:]), dtype=torch.float32, device=dev)
        idx = torch.empty(B, tout, dtype=torch.int64, device=dev) if return_index else None
        _lib.check(lib.efts_length_regulator_fwd(_ptr(x), _ptr(ds_eff), _ptr(il), _ptr(out_lens), B, T1, D,
                                                 tout, float(pad_value), _ptr(out), _ptr(idx), st))
    out = out.to(xs.dtype)
    return (out, idx) if return_index else out


# ---------------------------------------------------------------------------------------------------------------
# training slice (include/efts_b200.h, efts_resconv_train_*): ResConvBlock forward with saved activations + backward
class TrainContext:
    """A weight-less library context on one device for the training entry points, which take the caller's CURRENT
    fp32 weights on every call (they change every optimiser step, so nothing is prepacked on the host)."""

    def __init__(self, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("efts_b200 trains on CUDA (sm_100a) devices only; got %s" % (self.device,))
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        cfg = _lib.EftsConfig(1, 8, N_CHANNELS, 5, 1, 1, 1, 1, 3, 0.01, 0.5, 1.0, 0.1, 1, idx)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.efts_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self._ws = None

    def launch_count(self):
        return int(self.lib.efts_launch_count(self._h))

    def set_option(self, name, value):
        _lib.check(self.lib.efts_set_option(self._h, name.encode(), int(value)))

    def _workspace(self, B, T, k):
        n = int(self.lib.efts_resconv_train_workspace_bytes(self._h, B, T, k))
        if self._ws is None or self._ws.numel() < n:
            self._ws = None
            self._ws = torch.empty(n, dtype=torch.uint8, device=self.device)
        return self._ws

    def _check(self):
        flags = ctypes.c_int32(0)
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.efts_error_flags(self._h, st, ctypes.byref(flags)))
        if flags.value & 8:
            raise FloatingPointError(RANGE_MESSAGE)

    def resconv_fwd(self, x_btc, w_all, b_all):
        """x [B, T, 512], w_all [L, 512, 512, k], b_all [L, 512] -> (acts [L + 1, B, T, 512], us [L, B, T, 512])."""
        x = x_btc.detach().to(torch.float32).contiguous()
        w = w_all.detach().to(torch.float32).contiguous()
        b = b_all.detach().to(torch.float32).contiguous()
        B, T, C = x.shape
        L, k = w.shape[0], w.shape[3]
        if C != N_CHANNELS or tuple(w.shape[1:3]) != (C, C) or tuple(b.shape) != (L, C):
            raise RuntimeError("ResConv training kernels are built for %d channels; got x %s, weights %s" % (
                N_CHANNELS, tuple(x.shape), tuple(w.shape)))
        with torch.cuda.device(self.device):
            acts = torch.empty(L + 1, B, T, C, dtype=torch.float32, device=self.device)
            us = torch.empty(L, B, T, C, dtype=torch.float32, device=self.device)
            ws = self._workspace(B, T, k)
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.efts_resconv_train_fwd(self._h, _ptr(x), _ptr(w), _ptr(b), L, k, B, T, _ptr(acts), _ptr(us),
                                                       _ptr(ws), ws.numel(), st))
        return acts, us

    def resconv_bwd(self, grad_out, acts, us, w_all):
        g = grad_out.detach().to(torch.float32).contiguous()
        w = w_all.detach().to(torch.float32).contiguous()
        L, B, T, C = us.shape
        k = w.shape[3]
        with torch.cuda.device(self.device):
            gx = torch.empty(B, T, C, dtype=torch.float32, device=self.device)
            gw = torch.empty_like(w)
            gb = torch.empty(L, C, dtype=torch.float32, device=self.device)
            ws = self._workspace(B, T, k)
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.efts_resconv_train_bwd(self._h, _ptr(g), _ptr(acts), _ptr(us), _ptr(w), L, k, B, T, _ptr(gx),
                                                       _ptr(gw), _ptr(gb), _ptr(ws), ws.numel(), st))
            self._check()
        return gx, gw, gb


    # ---- duration predictor (layers/duration_predictor.py:57-88)
    def _dp_workspace(self, B, T, k):
        n = int(self.lib.efts_duration_train_workspace_bytes(self._h, B, T, k))
        if self._ws is None or self._ws.numel() < n:
            self._ws = None
            self._ws = torch.empty(n, dtype=torch.uint8, device=self.device)
        return self._ws

    def duration_fwd(self, x_btc, conv_w, conv_b, ln_g, ln_b, head_w, head_b, mask=None, keep=None):
        """x [B,T,512]; conv_w [L,512,512,k]; conv_b / ln_g / ln_b [L,512]; head_w [512]; head_b [1]; mask bool [B,T]
        (True = padded); keep [L,B,T,512] scaled dropout masks or None -> (out [B,T], acts, us)."""
        f = lambda t: t.detach().to(torch.float32).contiguous()
        x, conv_w, conv_b, ln_g, ln_b, head_w, head_b = (f(t) for t in (x_btc, conv_w, conv_b, ln_g, ln_b, head_w, head_b))
        B, T, C = x.shape
        L, k = conv_w.shape[0], conv_w.shape[3]
        if C != N_CHANNELS or tuple(conv_w.shape[1:3]) != (C, C) or head_w.numel() != C:
            raise RuntimeError("duration-predictor training kernels are built for %d channels" % N_CHANNELS)
        m = None if mask is None else mask.to(torch.uint8).contiguous()
        kp = None if keep is None else f(keep)
        with torch.cuda.device(self.device):
            acts = torch.empty(L + 1, B, T, C, dtype=torch.float32, device=self.device)
            us = torch.empty(L, B, T, C, dtype=torch.float32, device=self.device)
            out = torch.empty(B, T, dtype=torch.float32, device=self.device)
            ws = self._dp_workspace(B, T, k)
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.efts_duration_train_fwd(
                self._h, _ptr(x), _ptr(conv_w), _ptr(conv_b), _ptr(ln_g), _ptr(ln_b), _ptr(head_w), _ptr(head_b),
                _ptr(m) if m is not None else None, _ptr(kp) if kp is not None else None, L, k, B, T, _ptr(acts), _ptr(us),
                _ptr(out), _ptr(ws), ws.numel(), st))
        return out, acts, us

    def duration_bwd(self, grad_out, acts, us, conv_w, ln_g, head_w, mask=None, keep=None):
        f = lambda t: t.detach().to(torch.float32).contiguous()
        g, conv_w, ln_g, head_w = f(grad_out), f(conv_w), f(ln_g), f(head_w)
        L, B, T, C = us.shape
        k = conv_w.shape[3]
        m = None if mask is None else mask.to(torch.uint8).contiguous()
        kp = None if keep is None else f(keep)
        with torch.cuda.device(self.device):
            new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=self.device)
            gx, gw, gb, gg, gbeta, ghw, ghb = new(B, T, C), torch.empty_like(conv_w), new(L, C), new(L, C), new(L, C), new(C), new(1)
            ws = self._dp_workspace(B, T, k)
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.efts_duration_train_bwd(
                self._h, _ptr(g), _ptr(acts), _ptr(us), _ptr(conv_w), _ptr(ln_g), _ptr(head_w),
                _ptr(m) if m is not None else None, _ptr(kp) if kp is not None else None, L, k, B, T, _ptr(gx), _ptr(gw),
                _ptr(gb), _ptr(gg), _ptr(gbeta), _ptr(ghw), _ptr(ghb), _ptr(ws), ws.numel(), st))
            self._check()
        return gx, gw, gb, gg, gbeta, ghw, ghb

    # ---- criterion (losses/fastspeech_loss.py:54-67)
    def fastspeech_loss(self, before_outs, d_outs, ys, ds, ilens, olens, use_masking, use_mse, want_grads=True):
        """-> (losses fp32 [2] on the device: mel term, duration term; d mel term / d before_outs; d duration term /
        d d_outs) -- the gradients are None unless ``want_grads``."""
        f = lambda t: t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        mel, dur, ys, ds = f(before_outs), f(d_outs), f(ys), f(ds)
        il = ilens.detach().to(device=self.device, dtype=torch.int64).contiguous()
        ol = olens.detach().to(device=self.device, dtype=torch.int64).contiguous()
        B, T2, odim = mel.shape
        T1 = dur.shape[1]
        if ys.shape != mel.shape or ds.shape != dur.shape or il.numel() != B or ol.numel() != B:
            raise RuntimeError("FastSpeechLoss: shape mismatch (before_outs %s, ys %s, d_outs %s, ds %s)" % (
                tuple(mel.shape), tuple(ys.shape), tuple(dur.shape), tuple(ds.shape)))
        with torch.cuda.device(self.device):
            losses = torch.empty(2, dtype=torch.float32, device=self.device)
            gm = torch.empty_like(mel) if want_grads else None
            gd = torch.empty_like(dur) if want_grads else None
            n = int(self.lib.efts_fastspeech_loss_workspace_bytes(self._h))
            ws = torch.empty(n, dtype=torch.uint8, device=self.device)
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.efts_fastspeech_loss(
                self._h, _ptr(mel), _ptr(dur), _ptr(ys), _ptr(ds), _ptr(il), _ptr(ol), B, T1, T2, odim, int(use_masking),
                int(use_mse), _ptr(losses), _ptr(gm) if gm is not None else None, _ptr(gd) if gd is not None else None,
                _ptr(ws), ws.numel(), st))
        return losses, gm, gd

    def scale_by_scalar(self, x, scalar):
        """x * scalar with the 0-dim ``scalar`` read on the device."""
        x = x.contiguous()
        s = scalar.detach().to(device=self.device, dtype=torch.float32).reshape(1).contiguous()
        with torch.cuda.device(self.device):
            out = torch.empty_like(x)
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.efts_scale_by_scalar(self._h, _ptr(x), _ptr(s), x.numel(), _ptr(out), st))
        return out


_TRAIN_CONTEXTS = {}


def train_context(device):
    dev = torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    tc = _TRAIN_CONTEXTS.get(key)
    if tc is None:
        tc = _TRAIN_CONTEXTS[key] = TrainContext(dev)
    return tc


class ResConvStackFunction(torch.autograd.Function):
    """y = stack of ``x + lrelu(conv_k(x) + b)`` layers on channels-last activations, differentiable: forward and
    backward are library calls (efts_resconv_train_fwd / _bwd); ``w_all`` are the EFFECTIVE conv weights
    [L, C, C, k] (the weight-norm fold is left to torch so that weight_g / weight_v receive their gradients)."""

    @staticmethod
    def forward(ctx, x_btc, w_all, b_all):
        tc = train_context(x_btc.device)
        acts, us = tc.resconv_fwd(x_btc, w_all, b_all)
        ctx.save_for_backward(acts, us, w_all)
        ctx.tc = tc
        return acts[-1]

    @staticmethod
    def backward(ctx, grad_out):
        acts, us, w_all = ctx.saved_tensors
        gx, gw, gb = ctx.tc.resconv_bwd(grad_out, acts, us, w_all)
        return gx, gw, gb


class DurationPredictorFunction(torch.autograd.Function):
    """out [B, T] = Linear(LayerNorm(relu(conv(...)))) of layers/duration_predictor.py:69-86 (log domain, masked
    positions 0), differentiable: forward and backward are library calls (efts_duration_train_fwd / _bwd).  ``keep``
    holds the train-mode dropout masks (scaled by 1 / (1 - p)) drawn by the caller, or None."""

    @staticmethod
    def forward(ctx, x_btc, conv_w, conv_b, ln_g, ln_b, head_w, head_b, mask, keep):
        tc = train_context(x_btc.device)
        out, acts, us = tc.duration_fwd(x_btc, conv_w, conv_b, ln_g, ln_b, head_w, head_b, mask, keep)
        ctx.save_for_backward(acts, us, conv_w, ln_g, head_w)
        ctx.mask, ctx.keep, ctx.tc = mask, keep, tc
        return out

    @staticmethod
    def backward(ctx, grad_out):
        acts, us, conv_w, ln_g, head_w = ctx.saved_tensors
        gx, gw, gb, gg, gbeta, ghw, ghb = ctx.tc.duration_bwd(grad_out, acts, us, conv_w, ln_g, head_w, ctx.mask, ctx.keep)
        return gx, gw, gb, gg, gbeta, ghw.view_as(head_w), ghb, None, None


class FastSpeechLossFunction(torch.autograd.Function):
    """(mel term, duration term) of losses/fastspeech_loss.py:54-67 as 0-dim tensors, differentiable with respect to
    ``before_outs`` and ``d_outs``: one library call computes both terms and their element gradients; backward scales
    the saved gradients by the upstream scalars on the device."""

    @staticmethod
    def forward(ctx, before_outs, d_outs, ys, ds, ilens, olens, use_masking, use_mse):
        tc = train_context(before_outs.device)
        need = before_outs.requires_grad or d_outs.requires_grad
        losses, gm, gd = tc.fastspeech_loss(before_outs, d_outs, ys, ds, ilens, olens, use_masking, use_mse, need)
        if need:
            ctx.save_for_backward(gm, gd)
        ctx.tc = tc
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_mel, g_dur):
        gm, gd = ctx.saved_tensors
        return (ctx.tc.scale_by_scalar(gm, g_mel) if ctx.needs_input_grad[0] else None,
                ctx.tc.scale_by_scalar(gd, g_dur) if ctx.needs_input_grad[1] else None, None, None, None, None, None, None)
