"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, against the committed golden fixtures, and through size-independent properties.

Tolerances (stated by BASELINE.json's north_star): mel spectrogram <= 1e-4 abs vs the fp32
reference; length-regulator indices bit-exact.  The other returned tensors are held to the same
fp32-noise scale: reconst_alpha (values in [0,1]) <= 1e-4 abs; imv (values in [0, T1-1]) <= 2e-6
relative to its range, i.e. 4e-4 abs at T1 = 200 -- the reference's own fp32-vs-fp64 noise on imv is
7.5e-5 at that size.
"""
import os

import numpy as np
import pytest
import torch

from oracle import efts_oracle as orc
from tests.cases import make_forward_inputs, make_inference_inputs

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
MEL_TOL = 1e-4
RA_TOL = 1e-4


def imv_tol(t1):
    return max(1e-4, 2e-6 * t1)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def build_model(weights, dev):
    import efficient_tts_b200 as E
    m = E.EfficientTTSCNN(num_symbols=76, dropout_rate=0.0, use_masking=True, use_weighted_masking=False,
                          sigma=0.01)
    missing = m.load_state_dict(weights, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.eval().to(dev)


@pytest.fixture(scope="module")
def model(dev):
    return build_model(orc.make_weights(seed=1234), dev)


def run_both(model, dev, seed, t1, t2):
    w = orc.make_weights(seed=1234)
    text, tl, speech, sl = make_forward_inputs(seed, t1, t2)
    with torch.no_grad():
        ref = orc.forward(w, text, tl, speech, sl)
        out = model(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev),
                    speech_lengths=sl.to(dev))
    return ref, out, (text, tl, speech, sl)


def check_forward(ref, out, t1max):
    loss_r, stats_r, imv_r, ra_r, mel_r, _ = ref
    loss, stats, imv, ra, mel, _ = out
    d_mel = (mel.cpu() - mel_r).abs().max().item()
    d_ra = (ra.cpu() - ra_r).abs().max().item()
    d_imv = (imv.cpu() - imv_r).abs().max().item()
    print("max-abs: mel %.3e  reconst_alpha %.3e  imv %.3e  loss %.3e" %
          (d_mel, d_ra, d_imv, abs(float(loss) - float(loss_r))))
    assert d_mel <= MEL_TOL
    assert d_ra <= RA_TOL
    assert d_imv <= imv_tol(t1max)
    for k in ("loss", "mel_loss", "duration_loss"):
        assert abs(stats[k] - stats_r[k]) <= 1e-4 * max(1.0, abs(stats_r[k])), k
    assert abs(float(loss) - float(loss_r)) <= 1e-4 * max(1.0, abs(float(loss_r)))


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [
    dict(B=1, T=128, K=64, N=64, ntaps=1),
    dict(B=2, T=200, K=512, N=512, ntaps=1),
    dict(B=3, T=77, K=80, N=512, ntaps=1),          # K tail zero-filled by TMA (mel prenet shape)
    dict(B=2, T=300, K=512, N=80, ntaps=1),         # N tail (mel head shape)
    dict(B=2, T=260, K=512, N=512, ntaps=5),        # conv layer, tiles cross the utterance end
    dict(B=2, T=50, K=512, N=512, ntaps=3),         # duration-predictor conv
    dict(B=48, T=1100, K=512, N=512, ntaps=5),      # several tiles per persistent CTA (both epilogue groups)
])
@pytest.mark.parametrize("variant", ["pair", "single"])
def test_tap_gemm_against_float64(model, dev, shape, variant):
    """The tensor-core tap-GEMM (split-fp16, three products, main accumulator flushed every 2 k-blocks)
    reproduces an fp64 evaluation to fp32 level: within 8e-6, as CTA pairs and as single CTAs."""
    eng = model._get_engine()
    eng.set_option("pair", 1 if variant == "pair" else 0)
    try:
        g = torch.Generator().manual_seed(5)
        B, T, K, N, nt = shape["B"], shape["T"], shape["K"], shape["N"], shape["ntaps"]
        x = torch.randn(B, T, K, generator=g)
        w = torch.randn(nt, N, K, generator=g) / np.sqrt(K * nt)
        out = eng.tap_gemm(x.to(dev), w.to(dev), ntaps=nt, pad=(nt - 1) // 2).cpu()
        xp = torch.nn.functional.pad(x.double(), (0, 0, (nt - 1) // 2, (nt - 1) // 2))
        ref = sum(xp[:, j:j + T] @ w[j].double().T for j in range(nt))
        err = (out.double() - ref).abs().max().item()
        print("tap_gemm", shape, variant, "max-abs err %.3e" % err)
        assert err <= 8e-6
    finally:
        eng.set_option("pair", 1)


def test_batched_gemm_against_float64(model, dev):
    eng = model._get_engine()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 150, 512, generator=g)
    w = torch.randn(3, 40, 512, generator=g) / np.sqrt(512)
    out = eng.tap_gemm(x.to(dev), w.to(dev), batched=True).cpu()
    ref = torch.bmm(x.double(), w.double().transpose(1, 2))
    assert (out.double() - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("stack,name,n", [(0, "text_encoder", 5), (1, "mel_encoder", 3), (2, "decoder", 6)])
def test_conv_stack_matches_oracle(model, dev, stack, name, n):
    """ResConvBlock.forward (layers/efts_modules.py:77-79) incl. the non-inert padding region."""
    w = orc.make_weights(seed=1234)
    g = torch.Generator().manual_seed(stack)
    x = torch.randn(2, 512, 300, generator=g)
    with torch.no_grad():
        ref = orc.res_conv_stack(x, w, name, n)
    out = model._get_engine().conv_stack(stack, x.transpose(1, 2).contiguous().to(dev)).cpu().transpose(1, 2)
    err = (out - ref).abs().max().item()
    print(name, "max-abs err %.3e (|ref| max %.2f)" % (err, ref.abs().max().item()))
    assert err <= 1e-4


def test_resconvblock_module_matches_oracle(dev):
    """The stand-alone layer mirror, [B, C, T] in and out like the reference module."""
    from efficient_tts_b200.layers import ResConvBlock
    torch.manual_seed(3)
    blk = ResConvBlock(3).eval()
    w = {"blk." + k: v for k, v in blk.state_dict().items()}
    x = torch.randn(2, 512, 80)
    with torch.no_grad():
        ref = orc.res_conv_stack(x, w, "blk", 3)
    out = blk.to(dev)(x.to(dev)).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 1e-4


def test_duration_predictor_matches_golden_and_oracle(dev):
    from efficient_tts_b200.layers import DurationPredictor
    z = np.load(os.path.join(G, "dur_small.npz"))
    w = orc.make_weights(seed=1234, dur_bias=1.2)
    dp = DurationPredictor(idim=512, n_layers=2, n_chans=512, offset=1.0)
    dp.load_state_dict({k[len("duration_predictor."):]: v for k, v in w.items()
                        if k.startswith("duration_predictor.")})
    dp = dp.eval().to(dev)
    xs = torch.from_numpy(z["xs"]).to(dev)
    masks = (~orc.non_pad_mask(torch.from_numpy(z["lens"]))).to(dev)
    with torch.no_grad():                      # with gradients enabled forward() is the differentiable training path
        log_d = dp(xs, masks).cpu().numpy()
    log_d_grad = dp(xs, masks)                 # ... which returns the same values, attached to the autograd graph
    assert log_d_grad.requires_grad and np.abs(log_d_grad.detach().cpu().numpy() - log_d).max() <= 2e-5
    d_float = dp.inference(xs, None, to_round=False).cpu().numpy()
    d_round = dp.inference(xs, masks).cpu()
    np.testing.assert_allclose(log_d, z["log_d"], atol=1e-4, rtol=0)
    np.testing.assert_allclose(d_float, z["d_float"], atol=1e-4, rtol=1e-4)
    assert d_round.dtype == torch.int64
    flips = d_round.numpy() != z["d_round"]
    frac = np.abs((np.exp(z["log_d"]) - 1.0) % 1.0 - 0.5)
    assert (~flips | (frac < 1e-3)).all()


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["fwd_small", "fwd_pad_heavy"])
def test_forward_matches_golden(model, dev, name):
    """Whole forward() against outputs of the unmodified reference (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(G, name + ".npz"))
    text, tl, speech, sl = make_forward_inputs(int(z["seed"]), z["t1"], z["t2"])
    loss, stats, imv, ra, mel, sp = model(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev),
                                          speech_lengths=sl.to(dev))
    assert (mel.cpu().numpy() - z["mel_pred"]).__abs__().max() <= MEL_TOL
    assert (ra.cpu().numpy() - z["reconst_alpha"]).__abs__().max() <= RA_TOL
    assert (imv.cpu().numpy() - z["imv"]).__abs__().max() <= imv_tol(int(max(z["t1"])))
    np.testing.assert_allclose([stats["loss"], stats["mel_loss"], stats["duration_loss"]], z["stats"],
                               rtol=1e-4, atol=1e-4)
    # pad regions are exactly zero (models/efficient_tts.py:186,199)
    for b in range(len(z["t1"])):
        assert not mel[b, int(z["t2"][b]):].any()
        assert not ra[b, int(z["t1"][b]):].any() and not ra[b, :, int(z["t2"][b]):].any()
        assert not imv[b, int(z["t2"][b]):].any()


@pytest.mark.parametrize("case", [
    dict(seed=11, t1=[30, 22, 8, 1], t2=[200, 131, 47, 9]),          # ragged, a 1-token utterance
    dict(seed=12, t1=[129], t2=[1025]),                              # B=1, tile boundaries + 1
    dict(seed=13, t1=[64, 64, 50], t2=[384, 256, 384]),              # lengths on tile boundaries
    dict(seed=14, t1=[200, 50], t2=[1200, 300]),                     # C3-like extremes in one batch
])
def test_forward_matches_oracle(model, dev, case):
    ref, out, _ = run_both(model, dev, case["seed"], case["t1"], case["t2"])
    check_forward(ref, out, max(case["t1"]))


def test_forward_long_form_shape(model, dev):
    """C5-like shape (T1 = 300 needs two column tiles in the energy GEMM), B reduced to 2."""
    ref, out, _ = run_both(model, dev, 15, [300, 280], [2000, 1900])
    check_forward(ref, out, 300)


def test_padding_is_not_inert_but_batch_independent(model, dev):
    """An utterance's outputs depend on its padded length but not on what else is in
    the batch: run it alone and inside a batch with the same padded dims -> bitwise identical."""
    text, tl, speech, sl = make_forward_inputs(21, [40, 17, 25], [240, 100, 150])
    a = model(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    # same utterances in reversed order, same padded dims
    idx = torch.tensor([2, 1, 0])
    b = model(text=text[idx].to(dev), text_lengths=tl[idx].to(dev), speech=speech[idx].to(dev),
              speech_lengths=sl[idx].to(dev))
    for k in (2, 3, 4):
        assert torch.equal(a[k][idx], b[k])


def test_skip_pad_tiles_does_not_change_results(model, dev):
    text, tl, speech, sl = make_forward_inputs(22, [60, 20], [700, 130])
    args = dict(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    eng = model._get_engine()
    a = model(**args)
    eng.set_option("skip_pad_tiles", 0)
    try:
        b = model(**args)
    finally:
        eng.set_option("skip_pad_tiles", 1)
    for k in (2, 3, 4):
        assert torch.equal(a[k], b[k])
    assert abs(a[1]["loss"] - b[1]["loss"]) <= 1e-6 * max(1.0, abs(a[1]["loss"]))


def test_block_imv_kernels_equal_the_per_warp_kernels(model, dev):
    """The block-per-utterance scan / aligned-position kernels do the same arithmetic in the same order as the
    warp-per-row ones that serve rows too long for shared memory: every output bitwise equal."""
    text, tl, speech, sl = make_forward_inputs(23, [77, 200, 9, 130], [460, 1200, 50, 777])
    args = dict(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    eng = model._get_engine()
    a = model(**args)
    eng.set_option("imv_version", 1)
    try:
        b = model(**args)
    finally:
        eng.set_option("imv_version", 2)
    for k in (2, 3, 4):
        assert torch.equal(a[k], b[k])
    # pad tokens / pad frames of the returned matrix are exact zeros
    for bi in range(4):
        assert a[3][bi, int(tl[bi]):, :].abs().max().item() == 0.0 if int(tl[bi]) < a[3].shape[1] else True
        assert a[3][bi, :, int(sl[bi]):].abs().max().item() == 0.0 if int(sl[bi]) < a[3].shape[2] else True


def test_forward_rejects_what_the_reference_rejects(model, dev):
    text, tl, speech, sl = make_forward_inputs(23, [12, 9], [60, 40])
    with pytest.raises(RuntimeError):       # padded dim != max(lengths): nets_utils.py:148 broadcast error
        model(text=text.to(dev), text_lengths=(tl - 1).to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    bad = text.clone()
    bad[0, 0] = 76
    with pytest.raises(IndexError):         # embedding index out of range
        model(text=bad.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    long_sl = sl.clone()
    long_sl[1] = 5000                        # a length beyond the padded dim: clamped on the device, reported, no fault
    with pytest.raises(RuntimeError):
        model(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=long_sl.to(dev))
    neg_tl = tl.clone()
    neg_tl[1] = -3
    with pytest.raises(RuntimeError):
        model(text=text.to(dev), text_lengths=neg_tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    ok = model(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    assert torch.isfinite(ok[4]).all()       # the context is still healthy after the rejected calls
    with pytest.raises(RuntimeError):       # forward-only engine
        model.train()(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    model.eval()


# ------------------------------------------------------------------------------------------------
# BASELINE.json sizes against the CPU oracle (SURVEY.md 8d recipes): C2, C3 (seed 0 and two of the C4 shard
# seeds), C5.  Contract: every returned tensor is within 1e-4 abs of the fp32 reference -- or, where
# the reference's own fp32 rounding noise exceeds what any independent implementation can track, within 1e-4 of
# the reference evaluated in float64 AND closer to that float64 result than the fp32 reference itself is.
_PRECISION = {}


def _dump_precision():
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "precision_suite.json"), "w") as f:
            json.dump(_PRECISION, f, indent=1, sort_keys=True)
    except OSError:
        pass


FULL_SIZE_CASES = [("C2", 0), ("C3", 0), ("C3", 3), ("C3", 6), ("C5", 0)]


@pytest.mark.parametrize("name,seed", FULL_SIZE_CASES, ids=["%s_seed%d" % c for c in FULL_SIZE_CASES])
def test_forward_full_size_matches_oracle(model, dev, name, seed):
    import time
    from tests.cases import config_lengths
    w = orc.make_weights(seed=1234)
    t1, t2 = config_lengths(name, seed=seed)
    text, tl, speech, sl = make_forward_inputs(seed, t1, t2)
    out = model(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    t0 = time.time()
    with torch.no_grad():
        ref = orc.forward(w, text, tl, speech, sl)
    cpu_s = time.time() - t0
    r64 = None
    rec = {"cpu_oracle_seconds": cpu_s, "cpu_threads": torch.get_num_threads(), "B": len(t1)}
    failures = []
    for key, i, tol in (("mel", 4, MEL_TOL), ("reconst_alpha", 3, RA_TOL), ("imv", 2, 1e-4)):
        ours = out[i].cpu()
        d32 = (ours - ref[i]).abs().max().item()
        rec[key] = {"ours_vs_ref32_max": d32, "n_over_1e4": int(((ours - ref[i]).abs() > 1e-4).sum()),
                    "numel": ours.numel(), "rule": "direct"}
        if d32 <= tol:
            continue
        # the documented exception: the fp32 reference is itself further than the budget from the exact result
        if r64 is None:
            r64 = orc.forward_fp64(w, text, tl, speech, sl)
        ours64 = (ours.double() - r64[i]).abs().max().item()
        ref64 = (ref[i].double() - r64[i]).abs().max().item()
        rec[key].update(rule="fp64", ours_vs_fp64_max=ours64, ref32_vs_fp64_max=ref64)
        if not (ours64 <= tol and ours64 < ref64):
            failures.append("%s: %.3e from the fp32 reference, %.3e from fp64 (reference itself %.3e)" %
                            (key, d32, ours64, ref64))
    for k in ("loss", "mel_loss", "duration_loss"):
        rec[k] = [out[1][k], ref[1][k]]
        if abs(out[1][k] - ref[1][k]) > 1e-4 * max(1.0, abs(ref[1][k])):
            failures.append("%s: %r vs %r" % (k, out[1][k], ref[1][k]))
    _PRECISION["%s_seed%d" % (name, seed)] = rec
    _dump_precision()
    print(name, seed, rec)
    assert not failures, failures
    if name in ("C2", "C3"):
        # the bench configuration and its shards meet the north_star contract directly (no exception needed)
        assert rec["mel"]["rule"] == "direct" and rec["reconst_alpha"]["rule"] == "direct"


@pytest.mark.parametrize("name", ["C3", "C5"])
def test_forward_full_size_properties(model, dev, name):
    """BASELINE.json sizes (C3: B=256 mixed lengths padded to (200,1200); C5: B=32, T1=300, T2=2000), checked
    through properties that need no CPU run: exact zeros on padding, monotone IMV ending at T1-1, alignment
    columns that are probability vectors, losses that can be recomputed from the returned tensors, and
    bitwise batch independence (a slice of the batch with the same padded dims gives identical outputs)."""
    from tests.cases import config_lengths
    t1, t2 = config_lengths(name)
    text, tl, speech, sl = (t.to(dev) for t in make_forward_inputs(0, t1, t2))
    loss, stats, imv, ra, mel, _ = model(text=text, text_lengths=tl, speech=speech, speech_lengths=sl)
    B, T1, T2 = text.shape[0], text.shape[1], speech.shape[1]
    tm = torch.arange(T1, device=dev)[None] < tl[:, None]
    mm = torch.arange(T2, device=dev)[None] < sl[:, None]
    assert torch.isfinite(mel).all() and torch.isfinite(ra).all() and torch.isfinite(imv).all()
    assert not mel[~mm].any() and not imv[~mm].any()
    assert not ra[~(tm[:, :, None] & mm[:, None, :])].any()
    # imv_generator (models/efficient_tts.py:314-323): cumulative sum of ReLUs, normalised to T1_b - 1
    d = imv[:, 1:] - imv[:, :-1]
    assert bool((d[mm[:, 1:]] >= 0).all())
    last = imv.gather(1, (sl - 1)[:, None]).squeeze(1)
    assert torch.allclose(last, (tl - 1).float(), rtol=0, atol=1e-3)
    # reconstruct_align_from_aligned_position (:366-375): softmax over tokens on every valid frame
    col = ra.sum(1)
    assert torch.allclose(col[mm], torch.ones_like(col[mm]), rtol=0, atol=2e-5)
    assert float(ra.min()) >= 0.0
    # FastSpeechLoss (losses/fastspeech_loss.py:54-67): the mel term is recomputable from the outputs
    mel_loss = ((mel - speech) ** 2)[mm].double().mean().item()
    assert abs(stats["mel_loss"] - mel_loss) <= 1e-5 * max(1.0, mel_loss)
    assert abs(stats["loss"] - (stats["mel_loss"] + stats["duration_loss"])) <= 1e-5 * max(1.0, stats["loss"])
    # batch independence: rows 3..10 alone, padded to the same (T1, T2)
    idx = slice(3, 11)
    sub = model.forward_shard(text[idx], tl[idx], speech[idx], sl[idx])
    assert torch.equal(sub[0], imv[idx]) and torch.equal(sub[1], ra[idx]) and torch.equal(sub[2], mel[idx])


def test_forward_without_loss_masking(dev):
    """FastSpeechLoss(use_masking=False) (the reference's constructor default): means over the padded batch."""
    import efficient_tts_b200 as E
    w = orc.make_weights(seed=1234)
    m = E.EfficientTTSCNN(num_symbols=76, dropout_rate=0.0, use_masking=False, sigma=0.01)
    m.load_state_dict(w)
    m = m.eval().to(dev)
    text, tl, speech, sl = make_forward_inputs(24, [20, 11], [120, 70])
    with torch.no_grad():
        ref = orc.forward(w, text, tl, speech, sl, use_masking=False)
    out = m(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    for k in ("loss", "mel_loss", "duration_loss"):
        assert abs(out[1][k] - ref[1][k]) <= 1e-4 * max(1.0, abs(ref[1][k])), k
    assert (out[4].cpu() - ref[4]).abs().max().item() <= MEL_TOL


def test_forward_on_a_side_stream_and_with_long_text(model, dev):
    """Caller's stream is honoured; T1 > 512 takes the generic reconstruction kernel."""
    ref, _, (text, tl, speech, sl) = run_both(model, dev, 25, [530, 300], [700, 400])
    s = torch.cuda.Stream(dev)
    args = [t.to(dev) for t in (text, tl, speech, sl)]
    torch.cuda.synchronize(dev)
    with torch.cuda.stream(s):
        out = model(text=args[0], text_lengths=args[1], speech=args[2], speech_lengths=args[3])
    s.synchronize()
    check_forward(ref, out, 530)


def test_shard_keeps_global_padded_dims(model, dev):
    """forward_shard: a data-parallel shard whose own max length is below the global padded dims gives the
    same per-utterance outputs as the full batch, and its loss partials add up."""
    text, tl, speech, sl = make_forward_inputs(26, [33, 12, 25, 7], [210, 80, 160, 50])
    d = [t.to(dev) for t in (text, tl, speech, sl)]
    full = model.forward_shard(*d)
    parts = [model.forward_shard(*(t[lo:hi] for t in d)) for lo, hi in ((0, 2), (2, 4))]
    for k in range(3):
        assert torch.equal(torch.cat([p[k] for p in parts]), full[k])
    s = sum(p[3][3:7].cpu().double() for p in parts)
    f = full[3][3:7].cpu().double()
    assert torch.allclose(s, f, rtol=1e-6)
    with pytest.raises(RuntimeError):          # the reference-style call does check max(lengths) == padded dim
        model(text=d[0][2:], text_lengths=d[1][2:], speech=d[2][2:], speech_lengths=d[3][2:])


def test_fp16_range_violation_is_reported(dev):
    """Activations beyond the fp16 operand range are reported, not turned into inf/NaN silently."""
    w = orc.make_weights(seed=1234)
    w["text_embedding_table.weight"] = w["text_embedding_table.weight"] * 3e4
    m = build_model(w, dev)
    text, tl, speech, sl = make_forward_inputs(27, [12, 9], [60, 40])
    with pytest.raises(FloatingPointError):
        m(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    with pytest.raises(FloatingPointError):
        m.inference(text[:1, :12].to(dev))
    # NaN / inf in the inputs trip the same check (a float max would drop the NaN)
    m2 = build_model(orc.make_weights(seed=1234), dev)
    bad = speech.clone()
    bad[0, 3, 7] = float("nan")
    with pytest.raises(FloatingPointError):
        m2(text=text.to(dev), text_lengths=tl.to(dev), speech=bad.to(dev), speech_lengths=sl.to(dev))


# ------------------------------------------------------------------------------------------------
def test_inference_matches_golden(dev):
    z = np.load(os.path.join(G, "inf_small.npz"))
    w = orc.make_weights(seed=1234, dur_bias=float(z["dur_bias"]), dur_weight_scale=0.05)
    m = build_model(w, dev)
    text = make_inference_inputs(int(z["seed"]), int(z["t1"]))
    mel, ra = m.inference(text.to(dev))
    assert tuple(mel.shape) == z["mel_pred"].shape and tuple(ra.shape) == z["reconst_alpha"].shape
    assert np.abs(mel.cpu().numpy() - z["mel_pred"]).max() <= MEL_TOL
    assert np.abs(ra.cpu().numpy() - z["reconst_alpha"]).max() <= RA_TOL
    # the reference's loading order (bin/inference.py:77-83): load_state_dict, remove_weight_norm()
    # on the CPU, then .eval().to(device) -- folded weights give bit-identical results
    import efficient_tts_b200 as E
    m2 = E.EfficientTTSCNN(num_symbols=76, dropout_rate=0.0, use_masking=True, sigma=0.01)
    m2.load_state_dict(w)
    m2.remove_weight_norm()
    assert "decoder.layers.0.conv.0.weight" in m2.state_dict()
    assert "decoder.layers.0.conv.0.weight_g" not in m2.state_dict()
    m2 = m2.eval().to(dev)
    mel2, ra2 = m2.inference(text.to(dev))
    assert torch.equal(mel, mel2) and torch.equal(ra, ra2)


def test_inference_c1_shape_matches_oracle(dev):
    """BASELINE configs[0]: batch 1, 64 phonemes -> ~80 x 512 mel."""
    w = orc.make_weights(seed=1234, dur_bias=float(np.log(9.0)), dur_weight_scale=0.05)
    m = build_model(w, dev)
    text = make_inference_inputs(0, 64)
    with torch.no_grad():
        (mel_r, ra_r), inter = orc.inference(w, text, return_intermediates=True)
    e_last = float(inter["e"][0, -1])
    assert abs(e_last - np.floor(e_last) - 0.5) > 0.01, "T2 rounding margin too small for a stable test"
    mel, ra = m.inference(text.to(dev))
    assert mel.shape == mel_r.shape and 400 < mel.shape[1] < 640
    assert (mel.cpu() - mel_r).abs().max().item() <= MEL_TOL
    assert (ra.cpu() - ra_r).abs().max().item() <= RA_TOL
    with pytest.raises(RuntimeError):        # B > 1: the reference's .item() at :361 raises
        m.inference(torch.zeros(2, 5, dtype=torch.long, device=dev))


def test_resident_stack_and_split_reduction_are_bitwise_the_per_layer_result(dev):
    """B = 1 synthesis runs as two resident layer-stack kernels (stack_sm100.cuh: grid barriers between layers, every
    conv's reduction split into one work item per accumulation chunk and summed in chunk order).  Same bits as one
    launch per layer with the split reduction (splitk_reduce_kernel), and as the unsplit kernels."""
    w1 = orc.make_weights(seed=1234, dur_bias=1.7917594692, dur_weight_scale=0.05)
    m = build_model(w1, dev)
    eng = m._get_engine()
    modes = {"stack": dict(stack=1, split_k=1), "per_layer_split": dict(stack=0, split_k=1),
             "per_layer_unsplit": dict(stack=0, split_k=0)}
    for n_tok in (7, 40, 64, 129, 200):
        t = make_inference_inputs(11 + n_tok, n_tok).to(dev)
        res, launches = {}, {}
        try:
            for name, opts in modes.items():
                for k, v in opts.items():
                    eng.set_option(k, v)
                n0 = eng.launch_count()
                res[name] = m.inference(t)
                launches[name] = eng.launch_count() - n0
        finally:
            eng.set_option("stack", 1)
            eng.set_option("split_k", 1)
        # the paths really differ: 4 launches (stack, reconstruct, expand, stack) against one or two per layer
        assert launches["stack"] <= 5 < launches["per_layer_unsplit"] < launches["per_layer_split"], launches
        for name in ("per_layer_split", "per_layer_unsplit"):
            assert res[name][0].shape == res["stack"][0].shape
            assert torch.equal(res[name][0], res["stack"][0]) and torch.equal(res[name][1], res["stack"][1]), name
    # repeated runs of the resident kernel are deterministic (a missing grid barrier would show up as a race)
    t = make_inference_inputs(3, 64).to(dev)
    first = m.inference(t)
    for _ in range(10):
        again = m.inference(t)
        assert torch.equal(first[0], again[0]) and torch.equal(first[1], again[1])


def test_long_synthesis_leaves_the_resident_stack_and_still_matches_the_oracle(dev):
    """The resident stack holds its partial planes in a 16 MB scratch (T2 <= 2048 frames at 4 chunks); a longer
    synthesis takes one launch per layer.  Both sides of the limit against the oracle, and the launch counts show
    which path ran."""
    w = orc.make_weights(seed=1234, dur_bias=float(np.log(24.0)), dur_weight_scale=0.05)
    m = build_model(w, dev)
    eng = m._get_engine()
    for n_tok, stack_expected in ((80, True), (100, False)):
        t = make_inference_inputs(70 + n_tok, n_tok)
        with torch.no_grad():
            rmel, rra = orc.inference(w, t)
        n0 = eng.launch_count()
        mel, ra = m.inference(t.to(dev))
        launches = eng.launch_count() - n0
        assert mel.shape == rmel.shape and (mel.shape[1] <= 2048) == stack_expected, mel.shape
        assert (launches <= 6) == stack_expected, launches       # 4 launches with the stack, 10+ with one launch per layer
        assert (mel.cpu() - rmel).abs().max().item() <= MEL_TOL
        assert (ra.cpu() - rra).abs().max().item() <= RA_TOL


def test_programmatic_dependent_launch_does_not_change_results(model, dev):
    """GEMM launches carry the programmatic-stream-serialization attribute (their prologue overlaps the previous
    kernel's tail, `griddepcontrol.wait` guards the first touch of activations): bitwise the serialized result,
    over repeated runs (a missing wait would show up as a race)."""
    text, tl, speech, sl = make_forward_inputs(31, [33, 90, 12], [200, 540, 70])
    args = dict(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
    eng = model._get_engine()
    eng.set_option("pdl", 0)
    try:
        ref = model(**args)
    finally:
        eng.set_option("pdl", 1)
    for _ in range(5):
        out = model(**args)
        for k in (2, 3, 4):
            assert torch.equal(out[k], ref[k])
    w1 = orc.make_weights(seed=1234, dur_bias=1.7917594692, dur_weight_scale=0.05)
    m1 = build_model(w1, dev)
    e1 = m1._get_engine()
    t = make_inference_inputs(3, 48).to(dev)
    e1.set_option("pdl", 0)
    try:
        mel0, ra0 = m1.inference(t)
    finally:
        e1.set_option("pdl", 1)
    for _ in range(5):
        mel1, ra1 = m1.inference(t)
        assert torch.equal(mel0, mel1) and torch.equal(ra0, ra1)


def test_helper_methods_match_oracle(model, dev):
    """The reference's public stage methods (models/efficient_tts.py:287-398) called one by one."""
    w = orc.make_weights(seed=1234)
    text, tl, speech, sl = make_forward_inputs(51, [30, 18, 7], [190, 120, 45])
    with torch.no_grad():
        _, inter = orc.forward(w, text, tl, speech, sl, return_intermediates=True)
        text_mask, mel_mask = orc.non_pad_mask(tl), orc.non_pad_mask(sl)
        p_r = orc.index_vector(text_mask)
        alpha_r = orc.attention_alpha(inter["mel_h"], inter["key"], text_mask)
        alpha_r = alpha_r.masked_fill(~(text_mask.unsqueeze(-1) & mel_mask.unsqueeze(1)), 0.0)
        imv_r = orc.imv_from_alpha(alpha_r, p_r, mel_mask, tl)
        e_r = orc.aligned_positions(imv_r, p_r, mel_mask, text_mask, 0.5)
        ra_r = orc.reconstruct_alignment(e_r, 0.01, mel_mask, text_mask)
    tmd, mmd = text_mask.to(dev), mel_mask.to(dev)
    p = model.generate_index_vector(tmd)
    assert torch.equal(p.cpu(), p_r)
    alpha = model.scaled_dot_product_attention(inter["mel_h"].to(dev), inter["key"].to(dev), tmd)
    assert alpha.shape == alpha_r.shape
    live = mel_mask.unsqueeze(1).expand_as(alpha_r)
    assert (alpha.cpu() - orc.attention_alpha(inter["mel_h"], inter["key"], text_mask))[live].abs().max().item() <= 1e-5
    imv = model.imv_generator(alpha_r.to(dev), p, mmd, tl.to(dev))
    assert (imv.cpu() - imv_r).abs().max().item() <= 1e-4
    e = model.get_aligned_positions(imv_r.to(dev), p, mmd, tmd, sigma=0.5)
    # e spans [0, T2): held to the distance of the fp32 reference from its own float64 evaluation (a 190-term
    # softmax expectation), both against the reference and against float64
    with orc.float_is_double(), torch.no_grad():
        e64 = orc.aligned_positions(imv_r.double(), p_r.double(), mel_mask, text_mask, 0.5)
    ref_noise = (e_r.double() - e64).abs().max().item()
    d_e32, d_e64 = (e.cpu() - e_r).abs().max().item(), (e.cpu().double() - e64).abs().max().item()
    print("aligned positions: ours-ref32 %.3e  ours-fp64 %.3e  ref32-fp64 %.3e" % (d_e32, d_e64, ref_noise))
    assert d_e64 <= max(1e-4, 2.0 * ref_noise) and d_e32 <= max(1e-4, 3.0 * ref_noise)
    ra = model.reconstruct_align_from_aligned_position(e_r.to(dev), delta=0.01, mel_mask=mmd, text_mask=tmd)
    ra_ref = ra_r.masked_fill(~(text_mask.unsqueeze(-1) & mel_mask.unsqueeze(1)), 0.0)
    assert (ra.cpu() - ra_ref).abs().max().item() <= 1e-5
    # inference form: no masks, T2 = round(e[:, -1])
    e1 = torch.cumsum(torch.full((1, 9), 4.3), dim=1)
    with torch.no_grad():
        r1 = orc.reconstruct_alignment(e1, 0.01)
    g1 = model.reconstruct_align_from_aligned_position(e1.to(dev), delta=0.01)
    assert g1.shape == r1.shape and (g1.cpu() - r1).abs().max().item() <= 1e-5


def test_batched_inference_equals_per_utterance_inference(dev):
    """ragged batched synthesis.  Row b of the batch must be what the reference-shaped
    B = 1 call returns for that utterance alone (bitwise on this path, <= 1e-4 vs the oracle)."""
    w = orc.make_weights(seed=1234, dur_bias=float(np.log(7.0)), dur_weight_scale=0.05)
    m = build_model(w, dev)
    g = torch.Generator().manual_seed(41)
    lens = [37, 64, 5, 128, 129, 1]
    T1 = max(lens)
    text = torch.randint(0, 76, (len(lens), T1), generator=g)
    text = text * (torch.arange(T1)[None] < torch.tensor(lens)[:, None])       # id 0 on padding, any id works
    mel, mel_lens, ra = m.inference_batch(text.to(dev), torch.tensor(lens).to(dev))
    assert mel.shape[0] == len(lens) and mel.shape[1] == int(mel_lens.max()) and ra.shape[:2] == (len(lens), T1)
    for b, L in enumerate(lens):
        one_mel, one_ra = m.inference(text[b:b + 1, :L].to(dev))
        n = int(mel_lens[b])
        assert one_mel.shape[1] == n
        assert torch.equal(mel[b, :n], one_mel[0]), "utterance %d differs from its B=1 run" % b
        assert torch.equal(ra[b, :L, :n], one_ra[0])
        assert not mel[b, n:].any() and not ra[b, L:].any() and not ra[b, :, n:].any()
        with torch.no_grad():
            rmel, rra = orc.inference(w, text[b:b + 1, :L])
        assert rmel.shape[1] == n
        assert (one_mel.cpu() - rmel).abs().max().item() <= MEL_TOL
        assert (one_ra.cpu() - rra).abs().max().item() <= RA_TOL
    with pytest.raises(RuntimeError):
        m.inference_batch(text.to(dev), torch.tensor([0] + lens[1:]).to(dev))


# ------------------------------------------------------------------------------------------------
def test_length_regulator_bit_exact(dev):
    from efficient_tts_b200.engine import length_regulator
    from efficient_tts_b200.layers import LengthRegulator
    z = np.load(os.path.join(G, "lr_cases.npz"))
    lr = LengthRegulator()
    out = lr(torch.tensor([[[1.0], [2.0], [3.0]]], device=dev), torch.tensor([[1, 2, 3]], device=dev),
             torch.tensor([3], device=dev))
    assert np.array_equal(out.cpu().numpy(), z["doc_out"])                # docstring example :60-73
    xs, ilens = torch.from_numpy(z["xs"]).to(dev), torch.from_numpy(z["ilens"]).to(dev)
    ds = torch.from_numpy(z["ds_in"].copy()).to(dev)
    out = lr(xs, ds, ilens)
    assert np.array_equal(out.cpu().numpy(), z["out_a1"])
    assert np.array_equal(ds.cpu().numpy(), z["ds_after_a1"])             # in-place all-zero fix-up
    for alpha, key in ((1.3, "out_a13"), (0.5, "out_a05")):
        ds = torch.from_numpy(z["ds_in"].copy()).to(dev)
        assert np.array_equal(lr(xs, ds, ilens, alpha=alpha).cpu().numpy(), z[key])
        assert np.array_equal(ds.cpu().numpy(), z["ds_in"])
    out = LengthRegulator(pad_value=-9.0)(xs, torch.from_numpy(z["ds_in"].copy()).to(dev), ilens)
    assert np.array_equal(out.cpu().numpy(), z["out_pad9"])
    # indices against the oracle on a larger ragged batch, incl. empty rows and zero durations
    g = torch.Generator().manual_seed(31)
    B, T1, D = 37, 200, 512
    xs = torch.randn(B, T1, D, generator=g)
    ds = torch.randint(0, 12, (B, T1), generator=g)
    ilens = torch.randint(1, T1 + 1, (B,), generator=g)
    ds[5] = 0
    ref_out, ref_idx = orc.length_regulator(xs, ds.clone(), ilens)
    out, idx = length_regulator(xs.to(dev), ds.clone().to(dev), ilens.to(dev), return_index=True)
    assert torch.equal(idx.cpu(), ref_idx)                                # bit-exact index contract
    assert torch.equal(out.cpu(), ref_out)
    with pytest.raises(RuntimeError):
        bad = ds.clone()
        bad[0, 0] = -1
        length_regulator(xs.to(dev), bad.to(dev), ilens.to(dev))


def test_length_regulator_full_size_properties(dev):
    """BASELINE-size check through properties: frame counts, monotone indices, gather identity."""
    from efficient_tts_b200.engine import length_regulator
    g = torch.Generator(device="cpu").manual_seed(32)
    B, T1, D = 256, 200, 512
    xs = torch.randn(B, T1, D, generator=g).to(dev)
    ds = torch.randint(0, 13, (B, T1), generator=g).to(dev)
    ilens = torch.randint(50, T1 + 1, (B,), generator=g).to(dev)
    out, idx = length_regulator(xs, ds.clone(), ilens, return_index=True)
    valid = torch.arange(T1, device=dev)[None] < ilens[:, None]
    lens = (ds * valid).sum(1)
    assert out.shape[1] == int(lens.max())
    live = idx >= 0
    assert torch.equal(live.sum(1), lens)
    assert bool(((idx[:, 1:] >= idx[:, :-1]) | ~live[:, 1:]).all())       # sorted
    counts = torch.zeros(B, T1, dtype=torch.int64, device=dev).scatter_add_(1, idx.clamp(min=0), live.long())
    assert torch.equal(counts, ds * valid)                                # histogram of idx == durations
    gathered = torch.gather(xs, 1, idx.clamp(min=0)[..., None].expand(-1, -1, D)) * live[..., None]
    assert torch.equal(gathered, out)


# ------------------------------------------------------------------------------------------------
# HiFi-GAN V1 generator: waveform within 1e-4 abs of the fp32 reference (values in [-1, 1])
AUDIO_TOL = 1e-4


class _H(dict):
    __getattr__ = dict.__getitem__


@pytest.fixture(scope="module")
def vocoder(dev):
    from oracle import hifigan_oracle as hor
    from efficient_tts_b200.vocoder import Generator
    g = Generator(_H(hor.V1_CONFIG))
    res = g.load_state_dict(hor.make_weights(seed=4321), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return g.eval().to(dev)


@pytest.mark.parametrize("name", ["hifigan_small", "hifigan_batch"])
def test_vocoder_matches_golden_and_oracle(vocoder, dev, name):
    from oracle import hifigan_oracle as hor
    z = np.load(os.path.join(G, name + ".npz"))
    mel = hor.make_mel(int(z["seed"]), int(z["batch"]), int(z["frames"]))
    y = vocoder(mel.to(dev)).cpu()
    assert tuple(y.shape) == z["audio"].shape
    d = (y - torch.from_numpy(z["audio"])).abs().max().item()
    print("vocoder %s: max-abs vs reference fixture %.3e (signal max %.2f)" % (name, d, float(np.abs(z["audio"]).max())))
    assert d <= AUDIO_TOL


@pytest.mark.parametrize("batch,frames", [(1, 1), (1, 37), (3, 64), (1, 517)])
def test_vocoder_matches_oracle_across_lengths(vocoder, dev, batch, frames):
    """One frame (every layer is almost all zero padding), lengths that are not multiples of the 128-row tile,
    a batch, and the C1 length (517 frames -> 132 352 samples)."""
    from oracle import hifigan_oracle as hor
    mel = hor.make_mel(100 + frames, batch, frames)
    with torch.no_grad():
        ref = hor.generator_forward(hor.make_weights(seed=4321), mel)
    y = vocoder(mel.to(dev)).cpu()
    assert y.shape == ref.shape == (batch, 1, frames * 256)
    d = (y - ref).abs().max().item()
    print("vocoder B=%d T=%d: max-abs %.3e" % (batch, frames, d))
    assert d <= AUDIO_TOL


def test_vocoder_after_remove_weight_norm_and_batch_independence(dev):
    from oracle import hifigan_oracle as hor
    from efficient_tts_b200.vocoder import Generator
    g = Generator(_H(hor.V1_CONFIG))
    g.load_state_dict(hor.make_weights(seed=4321))
    g = g.eval().to(dev)
    mel = hor.make_mel(5, 2, 21).to(dev)
    a = g(mel)
    g.remove_weight_norm()
    b = g(mel)
    # remove_weight_norm() folds on the device, the engine folds weight_g / weight_v on the host: same formula,
    # different rounding of the norm
    assert (a - b).abs().max().item() <= 2e-6
    assert torch.equal(g(mel[1:2]), b[1:2])        # an utterance does not depend on its batch neighbours


def test_vocoder_grouped_packing_equals_plain_packing(dev):
    """The 32 / 64-channel layers run as super-tap GEMMs over [B, L / G, 128] views (pack_grouped); same products,
    differently grouped sums: within 2e-5 of the plain packing, and both within tolerance of the oracle."""
    from oracle import hifigan_oracle as hor
    from efficient_tts_b200.vocoder import Generator
    w = hor.make_weights(seed=4321)
    mel = hor.make_mel(55, 2, 33)
    with torch.no_grad():
        ref = hor.generator_forward(w, mel)
    outs = {}
    for grouped in (1, 0):
        g = Generator(_H(hor.V1_CONFIG))
        g.load_state_dict(w)
        g.set_engine_options(voc_group=grouped)
        g = g.eval().to(dev)
        n0 = g._get_engine().launch_count()
        outs[grouped] = g(mel.to(dev)).cpu()
        assert (outs[grouped] - ref).abs().max().item() <= AUDIO_TOL
    assert (outs[1] - outs[0]).abs().max().item() <= 2e-5


@pytest.mark.parametrize("name,cfg_name", [("hifigan_v2", "V2_CONFIG"), ("hifigan_v3", "V3_CONFIG")])
def test_vocoder_other_release_configs(dev, name, cfg_name):
    """The same Generator class builds the V2 (ResBlock1, 128 initial channels: stages of 64 / 32 / 16 / 8 channels,
    all on the grouped packing) and V3 (ResBlock2, three upsamplers, k = 7 with dilation 12 -> 200-row operand box)
    configurations of the HiFi-GAN release: reference fixtures and the oracle at a second length."""
    from oracle import hifigan_oracle as hor
    from efficient_tts_b200.vocoder import Generator
    cfg = getattr(hor, cfg_name)
    w = hor.make_weights(seed=4321, h=cfg)
    g = Generator(_H(cfg))
    res = g.load_state_dict(w, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    g = g.eval().to(dev)
    z = np.load(os.path.join(G, name + ".npz"))
    mel = hor.make_mel(int(z["seed"]), int(z["batch"]), int(z["frames"]))
    y = g(mel.to(dev)).cpu()
    d = (y - torch.from_numpy(z["audio"])).abs().max().item()
    print("vocoder %s: max-abs vs reference fixture %.3e" % (name, d))
    assert tuple(y.shape) == z["audio"].shape and d <= AUDIO_TOL
    mel2 = hor.make_mel(77, 1, 150)
    with torch.no_grad():
        ref = hor.generator_forward(w, mel2, cfg)
    d2 = (g(mel2.to(dev)).cpu() - ref).abs().max().item()
    print("vocoder %s T=150: max-abs %.3e" % (name, d2))
    assert d2 <= AUDIO_TOL
