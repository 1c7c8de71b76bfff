"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
product never touches the oracle, the drop-in mirrors the reference's parameter contract, and the
data-parallel split / reduce / gather logic works across 2 gloo ranks."""
import ctypes
import os
import re
import subprocess
import sys
import warnings

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "efts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(efts_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_the_header():
    from efficient_tts_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.SIGNATURES) == names          # the ctypes binding types exactly the header
    assert b"sm_100a" in lib.efts_version()


def test_no_device_is_an_error_not_a_fallback():
    """Without a GPU (or on the wrong architecture) the library refuses; nothing computes on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from efficient_tts_b200 import _lib
    lib = _lib.load()
    cfg = _lib.EftsConfig(76, 80, 512, 5, 5, 3, 6, 2, 3, 0.01, 0.5, 1.0, 0.1, 1, 0)
    h = ctypes.c_void_p()
    rc = lib.efts_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -3 and b"no CPU path" in lib.efts_last_error()
    import efficient_tts_b200 as E
    from efficient_tts_b200 import workloads as wl
    m = E.EfficientTTSCNN(**wl.MODEL_KWARGS).eval()
    text, tl, speech, sl = wl.make_forward_inputs(0, [5], [20])
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(text=text, text_lengths=tl, speech=speech, speech_lengths=sl)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.inference(text)


def test_create_rejects_unsupported_configs():
    from efficient_tts_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    bad = _lib.EftsConfig(76, 80, 384, 5, 5, 3, 6, 2, 3, 0.01, 0.5, 1.0, 0.1, 1, 0)       # n_channels
    assert lib.efts_create(ctypes.byref(bad), ctypes.byref(h)) == -2
    bad = _lib.EftsConfig(76, 80, 512, 7, 5, 3, 6, 2, 3, 0.01, 0.5, 1.0, 0.1, 1, 0)       # k_size
    assert lib.efts_create(ctypes.byref(bad), ctypes.byref(h)) == -2
    bad = _lib.EftsConfig(76, 80, 512, 5, 5, 3, 6, 2, 3, 0.01, 0.5, 1.0, 0.2, 1, 0)       # activation slope
    assert lib.efts_create(ctypes.byref(bad), ctypes.byref(h)) == -2
    assert lib.efts_create(None, ctypes.byref(h)) == -1
    import efficient_tts_b200 as E
    for kw in (dict(share_text_encoder_key_value=True), dict(use_mel_query_fc=True),
               dict(delta_e_method_1=False), dict(use_weighted_masking=True),
               dict(nonlinear_activation="ReLU", nonlinear_activation_params={})):
        with pytest.raises(NotImplementedError):
            E.EfficientTTSCNN(76, **kw)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "efficient_tts_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), "product file mentions the oracle: " + f
    code = "import sys; import efficient_tts_b200, efficient_tts_b200.engine, efficient_tts_b200.data_parallel, efficient_tts_b200.frontend, efficient_tts_b200.vocoder; " \
           "assert not any(m.startswith('oracle') for m in sys.modules); assert 'nntts' not in sys.modules"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_dropin_parameter_contract_matches_oracle_key_list():
    """state_dict keys/shapes == the reference's (the oracle's make_weights mirrors
    them and is pinned to the reference by tests/golden); load_state_dict(strict) round-trips;
    remove_weight_norm() folds to `.weight` exactly like torch's own fold."""
    from oracle import efts_oracle as orc
    import efficient_tts_b200 as E
    from efficient_tts_b200 import workloads as wl
    from efficient_tts_b200.engine import fold_state_dict
    warnings.filterwarnings("ignore")
    w = orc.make_weights(seed=1234)
    m = E.EfficientTTSCNN(**wl.MODEL_KWARGS)
    sd = m.state_dict()
    assert list(sd) == sorted(sd, key=list(sd).index) and set(sd) == set(w)
    for k in w:
        assert tuple(sd[k].shape) == tuple(w[k].shape), k
    m.load_state_dict(w, strict=True)
    folded = fold_state_dict(m.state_dict())
    m.remove_weight_norm()
    sd2 = m.state_dict()
    assert "decoder.layers.5.conv.0.weight" in sd2 and not any(k.endswith("weight_g") for k in sd2)
    for k, v in sd2.items():
        assert torch.equal(folded[k], v), k
        if k.endswith("conv.0.weight") and "duration" not in k:
            assert torch.equal(v, orc.conv_weight(w, k[: -len(".weight")]))
    m.apply_weight_norm()
    assert "decoder.layers.5.conv.0.weight_g" in m.state_dict()


def test_layer_mirrors_keep_reference_signatures():
    from efficient_tts_b200.layers import DurationPredictor, LengthRegulator, ResConvBlock
    blk = ResConvBlock(num_layers=2, n_channels=512, k_size=5, dropout_rate=0.0, use_weight_norm=True)
    assert list(blk.state_dict()) == ["layers.0.conv.0.bias", "layers.0.conv.0.weight_g", "layers.0.conv.0.weight_v",
                                      "layers.1.conv.0.bias", "layers.1.conv.0.weight_g", "layers.1.conv.0.weight_v"]
    dp = DurationPredictor(idim=512, n_layers=2, n_chans=512, kernel_size=3, dropout_rate=0.1, offset=1.0)
    assert "conv.1.2.weight" in dp.state_dict() and dp.state_dict()["linear.weight"].shape == (1, 512)
    assert dp.conv[0][2].eps == 1e-12                                    # layers/layer_norm.py:16
    assert LengthRegulator(pad_value=1.5).pad_value == 1.5
    with pytest.raises(NotImplementedError):
        DurationPredictor(idim=512, n_chans=512, spk_embed_dim=64, num_spks=4)


def test_workload_recipes():
    from efficient_tts_b200 import workloads as wl
    t1, t2 = wl.config_lengths("C3")
    assert len(t1) == 256 and sum(t1) == 32825 and max(t1) == 200 and min(t1) >= 50 and sum(t2) == 196950
    for seed in range(1, 8):
        a, b = wl.config_lengths("C3", seed=seed)
        assert max(a) == 200 and max(b) == 1200
    text, tl, speech, sl = wl.make_forward_inputs(0, [5, 3], [20, 9])
    assert text.shape == (2, 5) and speech.shape == (2, 20, 80)
    assert not text[1, 3:].any() and not speech[1, 9:].any()


# ------------------------------------------------------------------------------------------------
DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from efficient_tts_b200.data_parallel import DataParallelForward, shard_range, combine_loss_partials
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
B, T1, T2 = 5, 4, 6
g = torch.Generator().manual_seed(0)
text = torch.randint(0, 9, (B, T1), generator=g)
tl = torch.tensor([4, 2, 3, 1, 4]); sl = torch.tensor([6, 3, 5, 2, 6])
speech = torch.randn(B, T2, 2, generator=g)
def fake_forward(text, tl, speech, sl):       # stand-in with the device path's return contract
    mm = (torch.arange(T2)[None] < sl[:, None]).float()
    tm = (torch.arange(T1)[None] < tl[:, None]).float()
    mel = speech * 0.5 * mm[..., None]
    imv = mm * text.float().sum(1, keepdim=True)
    ra = tm[:, :, None] * mm[:, None, :]
    sq = (((mel - speech) ** 2) * mm[..., None]).sum()
    ab = (tm * text.float()).sum()
    scal = torch.stack([sq * 0, sq * 0, sq * 0, sq, mm.sum() * 2, ab, tm.sum(), sq * 0])
    return imv, ra, mel, scal
dp = DataParallelForward(fake_forward)
loss, stats, imv, ra, mel = dp(text, tl, speech, sl, gather_outputs=True)
ref = fake_forward(text, tl, speech, sl)
want = combine_loss_partials(ref[3][3:7].tolist())
assert abs(loss - want[0]) < 1e-6 and abs(stats["mel_loss"] - want[1]) < 1e-6, (loss, want)
assert torch.equal(imv, ref[0]) and torch.equal(ra, ref[1]) and torch.equal(mel, ref[2])
lo, hi = shard_range(B, rank, world)
_, _, imv_s, _, _ = dp(text, tl, speech, sl)             # outputs stay sharded by default
assert torch.equal(imv_s, ref[0][lo:hi])
# an error bit raised on ONE shard (token id out of range on rank 1's rows) stops every rank, like forward()
def flagged_forward(text, tl, speech, sl):
    out = fake_forward(text, tl, speech, sl)
    out[3][7] = 4.0 if rank == 1 else 0.0
    return out
try:
    DataParallelForward(flagged_forward)(text, tl, speech, sl)
    raise SystemExit("flag was not propagated on rank %%d" %% rank)
except IndexError:
    pass
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_data_parallel_split_reduce_gather_gloo_world2():
    from efficient_tts_b200.data_parallel import shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_range(2048, r, 8) for r in range(8)][3] == (768, 1024)
    port = 29500 + os.getpid() % 2000
    code = DP_WORKER % dict(root=ROOT, port=port)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
        assert "ok" in out


def test_bench_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--cpu-sample", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "mel_frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


# ---------------------------------------------------------------- HiFi-GAN generator mirror
def test_vocoder_mirror_keeps_the_reference_parameter_contract():
    """``efficient_tts_b200.vocoder.Generator(h)`` registers exactly the parameters of the reference Generator
    (vocoders/hifigan_model.py:97-118): the oracle's key list (pinned to the reference through the golden
    fixtures' strict load_state_dict) loads strictly, before and after remove_weight_norm()."""
    from oracle import hifigan_oracle as hor
    from efficient_tts_b200.vocoder import Generator

    class H(dict):
        __getattr__ = dict.__getitem__

    w = hor.make_weights(seed=4321)
    g = Generator(H(hor.V1_CONFIG))
    res = g.load_state_dict(w, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert sum(p.numel() for p in g.parameters()) == sum(v.numel() for v in w.values())
    g.remove_weight_norm()
    keys = set(g.state_dict().keys())
    assert "conv_pre.weight" in keys and "ups.0.weight" in keys and "resblocks.11.convs2.2.weight" in keys
    assert not any(k.endswith("weight_g") or k.endswith("weight_v") for k in keys)
    # no device, no result: the forward never falls back to the CPU
    with pytest.raises(RuntimeError):
        g.eval()(hor.make_mel(1, 1, 4))
    # the other two configurations the same class builds (ResBlock1 with 128 channels; ResBlock2)
    for cfg in (hor.V2_CONFIG, hor.V3_CONFIG):
        g2 = Generator(H(cfg))
        res = g2.load_state_dict(hor.make_weights(seed=4321, h=cfg), strict=True)
        assert not res.missing_keys and not res.unexpected_keys


def test_vocoder_create_rejects_unsupported_topologies():
    import ctypes
    from efficient_tts_b200 import _lib
    lib = _lib.load()
    cfg = _lib.EftsVocoderConfig()
    cfg.num_mels, cfg.upsample_initial_channel, cfg.num_upsamples, cfg.num_kernels = 80, 512, 1, 1
    cfg.upsample_rates[0], cfg.upsample_kernel_sizes[0] = 8, 12          # kernel != 2 * rate
    cfg.resblock_kernel_sizes[0] = 3
    for m in range(3):
        cfg.resblock_dilations[0][m] = 1
    cfg.resblock_type, cfg.num_dilations = 1, 3
    h = ctypes.c_void_p()
    assert lib.efts_vocoder_create(ctypes.byref(cfg), ctypes.byref(h)) == -2     # EFTS_ERR_UNSUPPORTED
    cfg.upsample_kernel_sizes[0] = 16
    cfg.resblock_kernel_sizes[0] = 13                                            # more than 11 taps
    assert lib.efts_vocoder_create(ctypes.byref(cfg), ctypes.byref(h)) == -2
    cfg.resblock_kernel_sizes[0] = 7
    cfg.resblock_dilations[0][1] = 13                                            # halo 78 rows > 72
    assert lib.efts_vocoder_create(ctypes.byref(cfg), ctypes.byref(h)) == -2
    cfg.resblock_dilations[0][1] = 1
    cfg.resblock_type = 3
    assert lib.efts_vocoder_create(ctypes.byref(cfg), ctypes.byref(h)) == -2


def test_vocoder_workload_recipe_equals_the_oracle_recipe():
    from oracle import hifigan_oracle as hor
    from efficient_tts_b200 import workloads as wl
    a, b = wl.vocoder_state_dict(4321), hor.make_weights(4321)
    assert a.keys() == b.keys()
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(wl.make_mel(3, 2, 5), hor.make_mel(3, 2, 5))


def test_load_hifigan_generator_mirror_reads_a_checkpoint(tmp_path):
    """Same steps as vocoders/hifigan_model.py:18-28 on a checkpoint file written in the reference's format."""
    import json
    from oracle import hifigan_oracle as hor
    from efficient_tts_b200.vocoder import load_hifigan_generator
    cfg = dict(hor.V2_CONFIG, sampling_rate=22050)
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    w = hor.make_weights(seed=1, h=cfg)
    torch.save({"generator": w}, tmp_path / "g_00000001")
    g = load_hifigan_generator("cpu", str(tmp_path / "config.json"), str(tmp_path / "g_00000001"))
    assert not g.training
    sd = g.state_dict()
    assert "conv_pre.weight" in sd and "conv_pre.weight_g" not in sd           # remove_weight_norm() ran
    assert torch.allclose(sd["conv_pre.weight"], hor.conv_weight(w, "conv_pre"), atol=1e-7)


def test_committed_bench_line_follows_the_contract():
    """The last bench line of the round (profiles/r01/bench_round1_final.json, produced on a B200 by
    tools/gpu_final.sh) carries every key the measurement contract names."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "profiles", "r01", "bench_round1_final.json")
    lines = [ln for ln in open(path).read().splitlines() if ln.strip()]
    assert len(lines) == 1                                   # exactly one line reached stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["config"]["workload"].startswith("C3") and "model" not in d["config"]
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("port", "reference")
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert abs(d["value"] - d["config"]["valid_frames_per_gpu"] * d["n_gpus"] / (d["ms_per_step"] * 1e-3)) < 1e-3 * d["value"]


# ---------------------------------------------------------------- generator weight re-arrangements (host-only hooks)
def _tap_gemm(x_blk, w_tkn, pad):
    """What the tap-GEMM computes: y[b, q] = sum_tau x[b, q + tau - pad] @ w[tau].T with zero rows outside [0, L)."""
    B, L, K = x_blk.shape
    S, N, _ = w_tkn.shape
    xp = torch.nn.functional.pad(x_blk.double(), (0, 0, pad, S - 1 - pad))
    return sum(xp[:, t:t + L] @ w_tkn[t].double().T for t in range(S))


@pytest.mark.parametrize("u,cin,cout", [(8, 16, 8), (2, 8, 8), (4, 8, 16)])
def test_transposed_conv_is_a_three_tap_polyphase_gemm(u, cin, cout):
    """efts_host_map_transposed (what pack_ups uploads): ConvTranspose1d(k = 2u, stride u, padding u/2) equals a
    3-tap GEMM whose N = u * Cout output columns are the u output phases of each input row
    (vocoders/hifigan_model.py:105-108,124)."""
    import ctypes
    from efficient_tts_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(u)
    k = 2 * u
    w = torch.randn(cin, cout, k, generator=g)
    out = torch.empty(3, u * cout, cin)
    assert lib.efts_host_map_transposed(w.data_ptr(), cin, cout, k, u, out.data_ptr()) == 0
    x = torch.randn(2, 7, cin, generator=g)                                   # [B, L, Cin] channels-last
    ref = torch.nn.functional.conv_transpose1d(x.transpose(1, 2).double(), w.double(), stride=u, padding=(k - u) // 2)
    y = _tap_gemm(x, out, 1).reshape(2, 7 * u, cout)                          # [B, L, u*Cout] IS [B, L*u, Cout]
    assert torch.allclose(y.transpose(1, 2), ref, atol=1e-10)
    assert lib.efts_host_map_transposed(w.data_ptr(), cin, cout, k + 2, u, out.data_ptr()) != 0   # k != 2u


@pytest.mark.parametrize("C,k,d,G", [(8, 3, 1, 4), (8, 11, 1, 4), (8, 7, 3, 4), (8, 11, 5, 4), (16, 7, 1, 2),
                                     (16, 3, 5, 2), (4, 5, 2, 8)])
def test_grouped_packing_is_the_dilated_conv(C, k, d, G):
    """efts_host_map_grouped (what pack_grouped uploads): a dilated Conv1d over [B, L, C] equals an undilated
    super-tap GEMM over the same buffer read as [B, L / G, G * C]; the tap count is 2 * floor((pad * d + G - 1) / G) + 1."""
    import ctypes
    from efficient_tts_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(C * k + d)
    w = torch.randn(C, C, k, generator=g)
    taps = ctypes.c_int32(0)
    assert lib.efts_host_map_grouped(w.data_ptr(), C, k, d, G, None, ctypes.byref(taps)) == 0
    S = taps.value
    assert S == 2 * (((k - 1) // 2 * d + G - 1) // G) + 1
    out = torch.empty(S, G * C, G * C)
    assert lib.efts_host_map_grouped(w.data_ptr(), C, k, d, G, out.data_ptr(), ctypes.byref(taps)) == 0
    L = 5 * G
    x = torch.randn(2, L, C, generator=g)
    ref = torch.nn.functional.conv1d(x.transpose(1, 2).double(), w.double(), dilation=d, padding=d * (k - 1) // 2)
    y = _tap_gemm(x.reshape(2, L // G, G * C), out, (S - 1) // 2).reshape(2, L, C)
    assert torch.allclose(y.transpose(1, 2), ref, atol=1e-10)


def test_frontend_host_logic_matches_the_checker():
    """The product's own filter bank and windowed DFT basis (efficient_tts_b200/frontend.py never imports test
    infrastructure): the filter bank equals the checker's restatement bit for bit, the DFT basis reproduces a float64
    FFT of a Hann-windowed frame, and the frame count follows the reference's padding rule."""
    from oracle import frontend_oracle as fo
    from efficient_tts_b200 import frontend as F
    assert np.array_equal(F.slaney_mel_basis(22050, 1024, 80, 0, 8000), fo.slaney_mel_basis())
    assert np.array_equal(F.slaney_mel_basis(16000, 512, 40, 50, None), fo.slaney_mel_basis(16000, 512, 40, 50, None))
    W = torch.from_numpy(F.windowed_dft_basis(1024, 256)).double().permute(0, 2, 1).reshape(1024, 1024)
    x = torch.randn(1024, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    ref = torch.fft.rfft(x * torch.hann_window(1024, dtype=torch.float64))
    out = W @ x
    assert (out[:513] - ref.real).abs().max() < 5e-6 and (out[513:] - ref.imag[1:512]).abs().max() < 5e-6
    from efficient_tts_b200 import _lib
    lib = _lib.load()
    for L in (256, 300, 1300, 22050, 256 * 40):
        # pure host arithmetic of the library (no device): frames of a padded utterance
        assert fo.num_frames(L) == 1 + (L + 768 - 1024) // 256
