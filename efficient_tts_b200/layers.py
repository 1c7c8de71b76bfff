"""Host-side mirrors of the ``nntts.layers`` modules on the EFTS-CNN path.

Same class names, constructor arguments, parameter names (``state_dict`` keys) and call signatures
as the reference (``/root/reference/nntts/layers/efts_modules.py``, ``duration_predictor.py``,
``layer_norm.py``, ``length_regulator.py``); the torch sub-modules only *hold* parameters -- the
arithmetic runs in ``libefts_b200.so``.  Forward-only: gradients are not produced.
"""
import torch

from . import engine as _engine


# Walking ``module.parameters()`` costs ~100 us for this model -- as much as a B = 1 synthesis takes on the GPU.  The
# list of Parameter objects is therefore cached per module and dropped whenever ANY module registers a parameter
# (weight-norm removal / application re-register ``weight``; plain ``.to()`` and ``load_state_dict`` keep the objects
# and show up as new ``data_ptr`` / ``_version`` values, which the fingerprint reads on every call).
_PARAM_GENERATION = [0]


def _on_parameter_registration(module, name, param):
    _PARAM_GENERATION[0] += 1


torch.nn.modules.module.register_module_parameter_registration_hook(_on_parameter_registration)


def _fingerprint(module):
    cache = module.__dict__.get("_efts_param_cache")
    if cache is None or cache[0] != _PARAM_GENERATION[0]:
        cache = (_PARAM_GENERATION[0], list(module.parameters()))
        module.__dict__["_efts_param_cache"] = cache
    return tuple([(p.data_ptr(), p._version) for p in cache[1]])


class _EngineOwner(torch.nn.Module):
    """Rebuilds the prepacked engine whenever parameters move or change."""

    def _engine_kwargs(self):
        raise NotImplementedError

    def _engine_state(self):
        raise NotImplementedError

    def _get_engine(self):
        fp = _fingerprint(self)
        eng = self.__dict__.get("_efts_engine")
        if eng is None or self.__dict__.get("_efts_fp") != fp:
            if eng is not None:
                eng.close()
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError(
                    "efficient_tts_b200 modules compute on a CUDA sm_100a device only (parameters are on "
                    "%s); there is no CPU path -- call .to('cuda')" % dev)
            eng = _engine.Engine(dev, self._engine_state(), **self._engine_kwargs())
            self.__dict__["_efts_engine"] = eng
            self.__dict__["_efts_fp"] = fp
        return eng

    def _require_eval(self):
        if self.training:
            raise RuntimeError("efficient_tts_b200 is a forward-only engine: call .eval() first (the "
                               "reference's train-mode dropout and autograd are out of scope)")


class LayerNorm(torch.nn.LayerNorm):
    """Parameter holder with the reference's signature (layers/layer_norm.py:6-30): eps 1e-12,
    normalisation over ``dim``.  Evaluated inside the duration-predictor kernels."""

    def __init__(self, nout, dim=-1):
        super().__init__(nout, eps=1e-12)
        self.dim = dim


class ResConv1d(torch.nn.Module):
    """Parameter holder for one residual layer (layers/efts_modules.py:19-51): ``conv.0`` is the
    Conv1d whose weights the tap-GEMM kernel consumes."""

    def __init__(self, n_channels=512, k_size=5, nonlinear_activation="LeakyReLU",
                 nonlinear_activation_params={"negative_slope": 0.1}, dropout_rate=0.1):
        super().__init__()
        if nonlinear_activation != "LeakyReLU" or float(nonlinear_activation_params.get("negative_slope", 0.01)) != 0.1:
            raise NotImplementedError("the sm_100a kernels implement LeakyReLU(0.1) only")
        mods = [torch.nn.Conv1d(n_channels, n_channels, kernel_size=k_size, padding=(k_size - 1) // 2),
                torch.nn.LeakyReLU(**nonlinear_activation_params)]
        if dropout_rate >= 1e-5:
            mods.append(torch.nn.Dropout(dropout_rate))     # identity in eval mode
        self.conv = torch.nn.Sequential(*mods)


def _apply_weight_norm(module):
    def fn(m):
        if isinstance(m, (torch.nn.Conv1d, torch.nn.ConvTranspose1d)):
            torch.nn.utils.weight_norm(m)
    module.apply(fn)


def _remove_weight_norm(module):
    def fn(m):
        try:
            torch.nn.utils.remove_weight_norm(m)
        except ValueError:
            return
    module.apply(fn)


class ResConvBlock(_EngineOwner):
    """``ResConvBlock(num_layers, ...).forward(x[B, C, T])`` (layers/efts_modules.py:54-79).  Forward-only through the
    prepacked engine under ``torch.no_grad()`` / ``.eval()``; with gradients enabled it runs the training slice
    (``engine.ResConvStackFunction``): same results, plus gradients for the input and every parameter."""

    def __init__(self, num_layers, n_channels=512, k_size=5, nonlinear_activation="LeakyReLU",
                 nonlinear_activation_params={"negative_slope": 0.1}, dropout_rate=0.1, use_weight_norm=True):
        super().__init__()
        self.num_layers = num_layers
        self.n_channels = n_channels
        self.k_size = k_size
        self.layers = torch.nn.Sequential(*[
            ResConv1d(n_channels, k_size, nonlinear_activation, nonlinear_activation_params, dropout_rate)
            for _ in range(num_layers)])
        if use_weight_norm:
            self.apply_weight_norm()

    def remove_weight_norm(self):
        _remove_weight_norm(self)

    def apply_weight_norm(self):
        _apply_weight_norm(self)

    # stand-alone use: a context holding this stack as its "decoder" (other weights are dummies)
    def _engine_kwargs(self):
        return dict(num_symbols=1, odim=8, n_channels=self.n_channels, k_size=self.k_size,
                    n_text_encoder_layer=1, n_mel_encoder_layer=1, n_decoder_layer=self.num_layers,
                    n_duration_layer=1)

    def _engine_state(self):
        C, k = self.n_channels, self.k_size
        sd = _dummy_state(C, k, 8, 1, 1, 1)
        for key in [k_ for k_ in sd if k_.startswith("decoder.")]:
            del sd[key]
        for key, v in self.state_dict().items():
            sd["decoder." + key] = v
        return sd

    def effective_weights(self):
        """Stacked effective conv weights [L, C, C, k] and biases [L, C], differentiable w.r.t. the module's
        parameters (``g * v / ||v||`` per output channel while weight norm is applied, layers/efts_modules.py:92-99)."""
        ws, bs = [], []
        for layer in self.layers:
            conv = layer.conv[0]
            if hasattr(conv, "weight_g"):
                ws.append(torch._weight_norm(conv.weight_v, conv.weight_g, 0))
            else:
                ws.append(conv.weight)
            bs.append(conv.bias)
        return torch.stack(ws), torch.stack(bs)

    def forward(self, x):
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if self.training or needs_grad:
            # training slice: forward that keeps what the backward needs, gradients from the library
            if self.training and any(isinstance(m, torch.nn.Dropout) and m.p > 0 for m in self.modules()):
                raise NotImplementedError("train-mode dropout is outside the B200 path (the production recipe sets "
                                          "dropout_rate=0.0, egs/lj/conf/*.yaml)")
            if x.device.type != "cuda":
                raise RuntimeError("efficient_tts_b200 modules compute on a CUDA sm_100a device only")
            w_all, b_all = self.effective_weights()
            y = _engine.ResConvStackFunction.apply(x.transpose(1, 2).contiguous(), w_all, b_all)
            return y.transpose(1, 2)
        y = self._get_engine().conv_stack(2, x.transpose(1, 2))
        return y.transpose(1, 2)


class DurationPredictor(_EngineOwner):
    """layers/duration_predictor.py:13-113 (speaker-embedding branch excluded: the EFTS model
    never enables it, models/efficient_tts.py:107-112)."""

    def __init__(self, idim, n_layers=2, n_chans=384, kernel_size=3, dropout_rate=0.1, offset=1.0,
                 num_spks=None, spk_embed_dim=None, spk_embed_integration_type="add"):
        super().__init__()
        if spk_embed_dim is not None or num_spks is not None:
            raise NotImplementedError("speaker-embedding duration predictor is outside the EFTS-CNN path")
        if idim != n_chans:
            raise NotImplementedError("the reference builds every conv with in_chans = n_chans "
                                      "(layers/duration_predictor.py:57); idim must equal n_chans")
        self.offset = offset
        self.n_layers = n_layers
        self.n_chans = n_chans
        self.kernel_size = kernel_size
        self.spk_embed_dim = None
        self.conv = torch.nn.ModuleList()
        for _ in range(n_layers):
            self.conv += [torch.nn.Sequential(
                torch.nn.Conv1d(n_chans, n_chans, kernel_size, stride=1, padding=(kernel_size - 1) // 2),
                torch.nn.ReLU(), LayerNorm(n_chans, dim=1), torch.nn.Dropout(dropout_rate))]
        self.linear = torch.nn.Linear(n_chans, 1)

    def _engine_kwargs(self):
        return dict(num_symbols=1, odim=8, n_channels=self.n_chans, k_size=1, n_text_encoder_layer=1,
                    n_mel_encoder_layer=1, n_decoder_layer=1, n_duration_layer=self.n_layers,
                    duration_kernel_size=self.kernel_size, duration_offset=self.offset)

    def _engine_state(self):
        sd = _dummy_state(self.n_chans, 1, 8, 1, 1, 1)
        for key, v in self.state_dict().items():
            sd["duration_predictor." + key] = v
        return sd

    def _run(self, xs, x_masks, mode):
        self._require_eval()
        out = self._get_engine().duration_predictor(xs, None, mode)
        if x_masks is not None:
            out = out.masked_fill(x_masks, 0.0)      # layers/duration_predictor.py:85-86
        return out

    def _train_forward(self, xs, x_masks):
        """Training slice: forward that keeps what the backward needs; gradients for every parameter and
        for ``xs`` come from the library (efts_duration_train_fwd / _bwd).  Train-mode dropout (the reference's default
        p = 0.1 here, layers/duration_predictor.py:62) uses masks drawn from torch's generator."""
        if xs.device.type != "cuda":
            raise RuntimeError("efficient_tts_b200 modules compute on a CUDA sm_100a device only")
        ws, bs, gs, betas = [], [], [], []
        for seq in self.conv:
            conv, ln = seq[0], seq[2]
            ws.append(torch._weight_norm(conv.weight_v, conv.weight_g, 0) if hasattr(conv, "weight_g") else conv.weight)
            bs.append(conv.bias)
            gs.append(ln.weight)
            betas.append(ln.bias)
        keep = None
        p = self.conv[0][3].p
        if self.training and p > 0:
            B, T, C = xs.shape
            keep = torch.nn.functional.dropout(torch.ones(self.n_layers, B, T, C, device=xs.device), p, True)
        return _engine.DurationPredictorFunction.apply(
            xs.contiguous(), torch.stack(ws), torch.stack(bs), torch.stack(gs), torch.stack(betas),
            self.linear.weight.reshape(-1), self.linear.bias, x_masks, keep)

    def forward(self, xs, x_masks=None, spembs=None):
        needs_grad = torch.is_grad_enabled() and (xs.requires_grad or any(p.requires_grad for p in self.parameters()))
        if self.training or needs_grad:
            return self._train_forward(xs, x_masks)
        return self._run(xs, x_masks, 0)

    def inference(self, xs, x_masks=None, spembs=None, to_round=True):
        return self._run(xs, x_masks, 2 if to_round else 1)


class LengthRegulator(torch.nn.Module):
    """layers/length_regulator.py:22-79: repeat token ``i`` ``ds[b, i]`` times, pad the batch."""

    def __init__(self, pad_value=0.0):
        super().__init__()
        self.pad_value = pad_value

    def forward(self, xs, ds, ilens, alpha=1.0):
        assert alpha > 0                               # layers/length_regulator.py:48
        return _engine.length_regulator(xs, ds, ilens, alpha=float(alpha), pad_value=float(self.pad_value))


def _dummy_state(C, k, odim, n_text, n_mel, n_dec, n_dur=0, dur_k=3):
    """Zero weights for the parts of a context a stand-alone layer does not use."""
    z = torch.zeros
    sd = {"text_embedding_table.weight": z(1, C)}
    for name, n in (("text_encoder", n_text), ("mel_encoder", n_mel), ("decoder", n_dec)):
        for i in range(n):
            sd["%s.layers.%d.conv.0.weight" % (name, i)] = z(C, C, k)
            sd["%s.layers.%d.conv.0.bias" % (name, i)] = z(C)
    for name in ("text_encoder_key", "text_encoder_value"):
        sd[name + ".weight"] = z(C, C)
        sd[name + ".bias"] = z(C)
    sd["mel_prenet.0.weight"] = z(C, odim)
    sd["mel_prenet.0.bias"] = z(C)
    sd["mel_output_layer.weight"] = z(odim, C)
    sd["mel_output_layer.bias"] = z(odim)
    for i in range(max(n_dur, 1)):
        sd["duration_predictor.conv.%d.0.weight" % i] = z(C, C, dur_k)
        sd["duration_predictor.conv.%d.0.bias" % i] = z(C)
        sd["duration_predictor.conv.%d.2.weight" % i] = z(C)
        sd["duration_predictor.conv.%d.2.bias" % i] = z(C)
    sd["duration_predictor.linear.weight"] = z(1, C)
    sd["duration_predictor.linear.bias"] = z(1)
    return sd
