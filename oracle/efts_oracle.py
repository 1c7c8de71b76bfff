"""CPU oracle for the EFTS-CNN forward path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference
``nntts.models.EfficientTTSCNN.forward()/.inference()`` and the ``nntts.layers``
duration predictor / length regulator.  It is the *checker* for the CUDA path in
``efficient_tts_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
package never does (``tests/test_host_logic.py::test_product_never_imports_the_oracle`` enforces that).

Where the arithmetic lives: the reference is pure Python on top of PyTorch, so its
numbers are "reference Python + this image's torch CPU kernels" (SURVEY.md 8c).  The
restatement therefore issues the same torch CPU ops in the same order -- written
functionally over a plain ``dict`` of tensors instead of ``nn.Module`` objects so it
has no dependency on ``/root/reference`` and can travel to the GPU box.

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, imported unmodified in the
build container by ``tests/golden/make_golden.py``; the resulting fixtures live in
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this file against them.

Reference citations are relative to ``/root/reference/nntts``.
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]

N_TEXT_LAYERS, N_MEL_LAYERS, N_DEC_LAYERS, N_DUR_LAYERS = 5, 3, 6, 2


# --------------------------------------------------------------------------- weights
def make_weights(num_symbols: int = 76, seed: int = 1234, odim: int = 80, chans: int = 512,
                 k_size: int = 5, dur_bias: Optional[float] = None,
                 dur_weight_scale: float = 1.0) -> Weights:
    """Deterministic random-init ``state_dict`` with the reference's key names and
    shapes (models/efficient_tts.py:57-112; key list in SURVEY.md 8b) and PyTorch's
    default init *scales* (uniform +-1/sqrt(fan_in) for conv/linear, N(0,1) embedding,
    ``weight_g = ||weight_v||`` as ``weight_norm`` sets it).  Layer-norm affine
    parameters are perturbed away from (1, 0) so the affine part is exercised."""
    g = torch.Generator().manual_seed(seed)

    def uni(*shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(*shape, generator=g) * 2 - 1) * b

    w: Weights = {}
    w["text_embedding_table.weight"] = torch.randn(num_symbols, chans, generator=g)
    for name, n in (("text_encoder", N_TEXT_LAYERS), ("mel_encoder", N_MEL_LAYERS),
                    ("decoder", N_DEC_LAYERS)):
        for i in range(n):
            v = uni(chans, chans, k_size, fan_in=chans * k_size)
            w[f"{name}.layers.{i}.conv.0.bias"] = uni(chans, fan_in=chans * k_size)
            w[f"{name}.layers.{i}.conv.0.weight_g"] = v.flatten(1).norm(dim=1).view(chans, 1, 1) \
                * (1.0 + 0.05 * torch.randn(chans, 1, 1, generator=g))
            w[f"{name}.layers.{i}.conv.0.weight_v"] = v
    for name in ("text_encoder_key", "text_encoder_value"):
        w[f"{name}.weight"] = uni(chans, chans, fan_in=chans)
        w[f"{name}.bias"] = uni(chans, fan_in=chans)
    w["mel_prenet.0.weight"] = uni(chans, odim, fan_in=odim)
    w["mel_prenet.0.bias"] = uni(chans, fan_in=odim)
    w["mel_output_layer.weight"] = uni(odim, chans, fan_in=chans)
    w["mel_output_layer.bias"] = uni(odim, fan_in=chans)
    for i in range(N_DUR_LAYERS):
        w[f"duration_predictor.conv.{i}.0.weight"] = uni(chans, chans, 3, fan_in=chans * 3)
        w[f"duration_predictor.conv.{i}.0.bias"] = uni(chans, fan_in=chans * 3)
        w[f"duration_predictor.conv.{i}.2.weight"] = 1.0 + 0.1 * torch.randn(chans, generator=g)
        w[f"duration_predictor.conv.{i}.2.bias"] = 0.1 * torch.randn(chans, generator=g)
    w["duration_predictor.linear.weight"] = uni(1, chans, fan_in=chans) * dur_weight_scale
    w["duration_predictor.linear.bias"] = uni(1, fan_in=chans)
    if dur_bias is not None:
        w["duration_predictor.linear.bias"] = torch.full((1,), float(dur_bias))
    return w


def conv_weight(w: Weights, prefix: str) -> torch.Tensor:
    """Effective conv weight: ``g * v / ||v||`` per output channel when the checkpoint
    still carries the weight-norm pair (layers/efts_modules.py:92-99, torch
    ``weight_norm`` dim 0), else the folded ``.weight`` left by ``remove_weight_norm``."""
    if prefix + ".weight" in w:
        return w[prefix + ".weight"]
    return torch._weight_norm(w[prefix + ".weight_v"], w[prefix + ".weight_g"], 0)


# --------------------------------------------------------------------------- masks
def non_pad_mask(lengths: torch.Tensor) -> torch.Tensor:
    """``mask[b, t] = t < lengths[b]`` with ``maxlen = max(lengths)``
    (utils/nets_utils.py:142-167 inverted at :254)."""
    lengths = lengths.to(torch.int64)
    steps = torch.arange(int(lengths.max()), dtype=torch.int64)
    return steps.unsqueeze(0) < lengths.unsqueeze(1)


# --------------------------------------------------------------------------- blocks
def res_conv_stack(x_bct: torch.Tensor, w: Weights, name: str, n_layers: int) -> torch.Tensor:
    """``x <- x + leaky_relu_0.1(conv1d_{k5,p2}(x) + b)`` per layer
    (layers/efts_modules.py:32-36,48-51,77-79); production ``dropout_rate=0`` so no dropout."""
    for i in range(n_layers):
        p = f"{name}.layers.{i}.conv.0"
        wt = conv_weight(w, p)
        y = F.conv1d(x_bct, wt, w[p + ".bias"], padding=(wt.shape[-1] - 1) // 2)
        x_bct = x_bct + F.leaky_relu(y, 0.1)
    return x_bct


def duration_predictor_log(xs_btc: torch.Tensor, w: Weights) -> torch.Tensor:
    """Eval-mode duration predictor up to the linear head, log domain
    (layers/duration_predictor.py:71-76; LayerNorm over channels with eps 1e-12,
    layers/layer_norm.py:16,30).  Returns [B, T]."""
    xs = xs_btc.transpose(1, -1)
    for i in range(N_DUR_LAYERS):
        p = f"duration_predictor.conv.{i}"
        xs = F.relu(F.conv1d(xs, w[p + ".0.weight"], w[p + ".0.bias"], padding=1))
        xs = F.layer_norm(xs.transpose(1, -1), (xs.shape[1],), w[p + ".2.weight"],
                          w[p + ".2.bias"], eps=1e-12).transpose(1, -1)
    return F.linear(xs.transpose(1, -1), w["duration_predictor.linear.weight"],
                    w["duration_predictor.linear.bias"]).squeeze(-1)


def duration_predictor_forward(xs, x_masks, w):
    """``DurationPredictor.forward`` (layers/duration_predictor.py:85-101)."""
    out = duration_predictor_log(xs, w)
    if x_masks is not None:
        out = out.masked_fill(x_masks, 0.0)
    return out


def duration_predictor_inference(xs, w, x_masks=None, to_round=True, offset=1.0):
    """``DurationPredictor.inference`` (layers/duration_predictor.py:78-88,103-113)."""
    out = duration_predictor_log(xs, w)
    if to_round:
        out = torch.clamp(torch.round(out.exp() - offset), min=0).long()
    else:
        out = torch.clamp(out.exp() - offset, min=0)
    if x_masks is not None:
        out = out.masked_fill(x_masks, 0.0)
    return out


# --------------------------------------------------------------------------- IMV chain
def attention_alpha(query, key, text_mask):
    """models/efficient_tts.py:377-398: softmax over T1 of QK^T / sqrt(D) with pad keys
    at -inf then zeroed; returned transposed to [B, T1, T2]."""
    D = key.size(-1)
    T2 = query.size(1)
    scores = torch.bmm(query, key.transpose(-2, -1)) / np.sqrt(float(D))
    kill = ~(text_mask.unsqueeze(1).repeat(1, T2, 1))
    scores = scores.masked_fill(kill, -float("inf"))
    alpha = torch.softmax(scores, dim=-1).masked_fill(kill, 0.0)
    return alpha.transpose(-2, -1)


def index_vector(text_mask):
    """models/efficient_tts.py:287-297."""
    B, T1 = text_mask.shape
    return torch.arange(0, T1).repeat(B, 1).float() * text_mask


def imv_from_alpha(alpha, p, mel_mask, text_lengths):
    """models/efficient_tts.py:299-324."""
    B = alpha.size(0)
    dummy = torch.bmm(alpha.transpose(1, 2), p.unsqueeze(-1)).squeeze(-1)
    d = torch.relu(dummy[:, 1:] - dummy[:, :-1])
    d = torch.cat([torch.zeros(B, 1).type_as(alpha), d], -1)
    imv = torch.cumsum(d, -1) * mel_mask.float()
    last, _ = torch.max(imv, dim=-1)
    last = torch.clamp(last, min=1e-8)
    return imv / last.unsqueeze(1) * (text_lengths.float().unsqueeze(-1) - 1)


def aligned_positions(imv, p, mel_mask, text_mask, sigma):
    """models/efficient_tts.py:326-345.  Returns [B, T1]."""
    en = -1 * ((imv.unsqueeze(1) - p.unsqueeze(-1)) ** 2) * sigma
    en = en.masked_fill(~(mel_mask.unsqueeze(1).repeat(1, en.size(1), 1)), -float("inf"))
    beta = torch.softmax(en, dim=2)
    q = torch.arange(0, mel_mask.size(-1)).unsqueeze(0).repeat(imv.size(0), 1).float()
    q = q * mel_mask.float()
    return (torch.bmm(beta, q.unsqueeze(-1)) * text_mask.unsqueeze(-1)).squeeze(-1)


def reconstruct_alignment(e, delta, mel_mask=None, text_mask=None):
    """models/efficient_tts.py:347-375 (``trim_e`` is only reachable with
    ``delta_e_method_1=False``, which is outside the production config)."""
    if mel_mask is None:
        max_length = torch.round(e[:, -1]).squeeze().item()
    else:
        max_length = mel_mask.size(-1)
    q = torch.arange(0, max_length).unsqueeze(0).repeat(e.size(0), 1).float()
    if mel_mask is not None:
        q = q * mel_mask.float()
    en = -1 * delta * (q.unsqueeze(1) - e.unsqueeze(-1)) ** 2
    if text_mask is not None:
        en = en.masked_fill(~(text_mask.unsqueeze(-1).repeat(1, 1, int(max_length))),
                            -float("inf"))
    return torch.softmax(en, dim=1)


# --------------------------------------------------------------------------- model
def forward(w: Weights, text, text_lengths, speech, speech_lengths, sigma=0.01, sigma_e=0.5,
            duration_offset=1.0, return_intermediates=False, use_masking=True):
    """Teacher-forced pass, models/efficient_tts.py:120-228, eval mode, production
    flags (``use_masking=True``, ``delta_e_method_1=True``, separate key/value, no
    mel-query fc).  Returns ``(loss, stats, imv, reconst_alpha, mel_pred, speech)``."""
    text_mask = non_pad_mask(text_lengths)
    mel_mask = non_pad_mask(speech_lengths)
    tm_mask = text_mask.unsqueeze(-1) & mel_mask.unsqueeze(1)

    emb = F.embedding(text, w["text_embedding_table.weight"]).transpose(1, 2)
    text_h = res_conv_stack(emb, w, "text_encoder", N_TEXT_LAYERS).transpose(1, 2)
    key = F.linear(text_h, w["text_encoder_key.weight"], w["text_encoder_key.bias"])
    value = F.linear(text_h, w["text_encoder_value.weight"], w["text_encoder_value.bias"])
    pad_tok = ~(text_mask.unsqueeze(-1).repeat(1, 1, text_h.size(2)))
    key = key.masked_fill(pad_tok, 0.0)
    value = value.masked_fill(pad_tok, 0.0)

    mel_h = F.leaky_relu(F.linear(speech, w["mel_prenet.0.weight"], w["mel_prenet.0.bias"]), 0.1)
    mel_h = res_conv_stack(mel_h.transpose(1, 2), w, "mel_encoder", N_MEL_LAYERS).transpose(1, 2)

    alpha = attention_alpha(mel_h, key, text_mask).masked_fill(~tm_mask, 0.0)
    p = index_vector(text_mask)
    imv = imv_from_alpha(alpha, p, mel_mask, text_lengths)
    e = aligned_positions(imv, p, mel_mask, text_mask, sigma_e)
    reconst = reconstruct_alignment(e, sigma, mel_mask, text_mask).masked_fill(~tm_mask, 0.0)

    expanded = torch.bmm(value.transpose(1, 2), reconst)
    expanded = expanded.masked_fill(~(mel_mask.unsqueeze(1).repeat(1, value.size(2), 1)), 0.0)
    dec = res_conv_stack(expanded, w, "decoder", N_DEC_LAYERS)
    mel_pred = F.linear(dec.transpose(1, 2), w["mel_output_layer.weight"],
                        w["mel_output_layer.bias"])
    mel_pred = mel_pred.masked_fill(~(mel_mask.unsqueeze(-1).repeat(1, 1, mel_pred.size(-1))), 0.0)

    delta_e = torch.cat([e[:, :1], e[:, 1:] - e[:, :-1]], dim=1)
    log_delta_e = torch.log(delta_e + duration_offset).masked_fill(~text_mask, 0.0)
    dur_pred = duration_predictor_forward(value, ~text_mask, w)

    # losses/fastspeech_loss.py:54-67: masked_select under use_masking=True, plain means otherwise
    if use_masking:
        mel_loss = F.mse_loss(mel_pred.masked_select(mel_mask.unsqueeze(-1)),
                              speech.masked_select(mel_mask.unsqueeze(-1)))
        dur_loss = F.l1_loss(dur_pred.masked_select(text_mask), log_delta_e.masked_select(text_mask))
    else:
        mel_loss = F.mse_loss(mel_pred, speech)
        dur_loss = F.l1_loss(dur_pred, log_delta_e)
    loss = mel_loss + dur_loss
    stats = dict(loss=loss.item(), mel_loss=mel_loss.item(), duration_loss=dur_loss.item())
    if return_intermediates:
        inter = dict(text_h=text_h, key=key, value=value, mel_h=mel_h, e=e, expanded=expanded,
                     dur_pred=dur_pred, log_delta_e=log_delta_e, dec=dec)
        return (loss, stats, imv, reconst, mel_pred, speech), inter
    return loss, stats, imv, reconst, mel_pred, speech


@contextlib.contextmanager
def float_is_double():
    """Inside the block ``Tensor.float()`` returns float64: the restatement's mask / index casts then stay in
    double when its inputs are double ("simplest fp64 restatement")."""
    orig = torch.Tensor.float
    torch.Tensor.float = lambda self, *a, **k: self.double()
    try:
        yield
    finally:
        torch.Tensor.float = orig


def forward_fp64(w: Weights, text, text_lengths, speech, speech_lengths, **kw):
    """The same teacher-forced pass evaluated in float64 (SURVEY.md 8c recipe: double weights and inputs, and
    ``Tensor.float`` patched to ``.double()`` for the duration of the call so the mask / index casts of
    models/efficient_tts.py:296,317,323,343,368 stay in double).  It is the tie-breaker where the reference's own
    fp32 rounding noise exceeds the parity budget: the distance of an implementation to this result
    is its own error, the distance of ``forward`` (fp32) to it is the reference's."""
    w64 = {k: v.double() for k, v in w.items()}
    with float_is_double(), torch.no_grad():
        return forward(w64, text, text_lengths, speech.double(), speech_lengths, **kw)


def inference(w: Weights, text, sigma=0.01, duration_offset=1.0, return_intermediates=False):
    """Free-running synthesis, models/efficient_tts.py:230-285 (B must be 1 because of
    the ``.item()`` at :361).  Returns ``(mel_pred[1,T2,odim], reconst_alpha[1,T1,T2])``."""
    emb = F.embedding(text, w["text_embedding_table.weight"]).transpose(1, 2)
    text_h = res_conv_stack(emb, w, "text_encoder", N_TEXT_LAYERS).transpose(1, 2)
    value = F.linear(text_h, w["text_encoder_value.weight"], w["text_encoder_value.bias"])
    delta_e = duration_predictor_inference(value, w, to_round=False, offset=duration_offset)
    e = torch.cumsum(delta_e, dim=1)
    reconst = reconstruct_alignment(e, sigma)
    expanded = torch.bmm(value.transpose(1, 2), reconst)
    dec = res_conv_stack(expanded, w, "decoder", N_DEC_LAYERS)
    mel = F.linear(dec.transpose(1, 2), w["mel_output_layer.weight"], w["mel_output_layer.bias"])
    if return_intermediates:
        return (mel, reconst), dict(value=value, delta_e=delta_e, e=e)
    return mel, reconst


# --------------------------------------------------------------------------- length regulator
def length_regulator(xs: torch.Tensor, ds: torch.Tensor, ilens: torch.Tensor, alpha: float = 1.0,
                     pad_value: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """``LengthRegulator.forward`` (layers/length_regulator.py:35-79, ``pad_list``
    utils/nets_utils.py:28-55), restated with numpy integer arithmetic.

    Returns ``(out[B, Tout, D], idx[B, Tout] int64)`` where ``idx[b, j]`` is the source
    token of output frame ``j`` (-1 on padding) -- the bit-exact index contract
    ``idx[b, j] = #{i : cumsum(ds[b])_i <= j}``.  Also mirrors the in-place all-zero
    fix-up on the caller's ``ds`` when ``alpha == 1.0`` (:76-78, the slices are views)."""
    assert alpha > 0
    if alpha != 1.0:
        ds = torch.round(ds.float() * alpha).long()
    B = xs.shape[0]
    rows = []
    for b in range(B):
        n = int(ilens[b])
        d = ds[b, :n]
        if int(d.sum()) == 0:
            d.fill_(1)                      # view: writes through to ``ds``
        dn = d.numpy().astype(np.int64)
        if (dn < 0).any():
            raise RuntimeError("negative duration")
        rows.append(np.repeat(np.arange(n, dtype=np.int64), dn))
    tout = max(len(r) for r in rows)
    idx = np.full((B, tout), -1, dtype=np.int64)
    out = xs.new_full((B, tout) + tuple(xs.shape[2:]), pad_value)
    for b, r in enumerate(rows):
        idx[b, :len(r)] = r
        out[b, :len(r)] = xs[b][torch.from_numpy(r)]
    return out, torch.from_numpy(idx)


# --------------------------------------------------------------------------- criterion (losses/fastspeech_loss.py)
def make_loss_inputs(seed, B, T1, T2, odim):
    """Seeded inputs of the criterion: (mel_pred [B,T2,odim], dur_pred [B,T1], ys, ds, ilens, olens); the longest
    utterance fills the padded dims (make_non_pad_mask builds its masks with maxlen = max(lengths))."""
    g = torch.Generator().manual_seed(int(seed))
    mel = torch.randn(B, T2, odim, generator=g)
    ys = torch.randn(B, T2, odim, generator=g)
    dur = torch.randn(B, T1, generator=g)
    ds = torch.randn(B, T1, generator=g)
    il = torch.randint(1, T1 + 1, (B,), generator=g)
    ol = torch.randint(1, T2 + 1, (B,), generator=g)
    il[0], ol[0] = T1, T2
    return mel, dur, ys, ds, il, ol


def fastspeech_loss(before_outs, d_outs, ys, ds, ilens, olens, use_masking=True, use_mse=True):
    """``FastSpeechLoss.forward`` with after_outs=None and use_weighted_masking=False (losses/fastspeech_loss.py:54-67):
    masked_select of the valid positions, then the mean criterion (MSE or L1 for the mel term, L1 for the durations)."""
    if use_masking:
        dm = non_pad_mask(ilens)
        d_outs, ds = d_outs.masked_select(dm), ds.masked_select(dm)
        om = non_pad_mask(olens).unsqueeze(-1)
        before_outs, ys = before_outs.masked_select(om), ys.masked_select(om)
    mel_loss = F.mse_loss(before_outs, ys) if use_mse else F.l1_loss(before_outs, ys)
    return mel_loss, F.l1_loss(d_outs, ds)
