#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``; it does not exist on the
GPU box):  ``python tests/golden/make_golden.py``

The reference ``nntts`` package is imported as-is, its ``EfficientTTSCNN`` is built with
the production kwargs (egs/lj/conf/efficient_tts_cnn_phnseq_noDropout.v1.yaml:17-22),
the deterministic weights of ``oracle.efts_oracle.make_weights`` are loaded through its
own ``load_state_dict``, and ``forward`` / ``inference`` / ``DurationPredictor`` /
``LengthRegulator`` are executed on seeded inputs.  Inputs that are cheap to regenerate
are stored as seeds; outputs are stored as arrays.  ``tests/test_oracle_golden.py`` pins
the oracle to these files; the GPU parity tests use them as a second opinion.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from oracle import efts_oracle as orc  # noqa: E402  (weights + input recipes only)
import nntts.models as ref_models  # noqa: E402
from nntts.layers.length_regulator import LengthRegulator  # noqa: E402
from nntts.utils.nets_utils import make_non_pad_mask, make_pad_mask  # noqa: E402


def ref_model(weights):
    net = ref_models.EfficientTTSCNN(num_symbols=76, dropout_rate=0.0, use_masking=True,
                                     use_weighted_masking=False, sigma=0.01)
    missing = net.load_state_dict(weights, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net.eval()


def forward_case(name, seed, t1, t2):
    from tests.cases import make_forward_inputs
    w = orc.make_weights(seed=1234)
    text, tl, speech, sl = make_forward_inputs(seed, t1, t2)
    net = ref_model(w)
    with torch.no_grad():
        loss, stats, imv, ra, mel, _ = net(text=text, text_lengths=tl, speech=speech,
                                           speech_lengths=sl)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, t1=np.array(t1),
                        t2=np.array(t2), loss=loss.numpy(),
                        stats=np.array([stats["loss"], stats["mel_loss"], stats["duration_loss"]]),
                        imv=imv.numpy(), reconst_alpha=ra.numpy(), mel_pred=mel.numpy())
    print(name, "loss", float(loss), "mel", tuple(mel.shape))


def inference_case(name, seed, t1, dur_bias):
    from tests.cases import make_inference_inputs
    w = orc.make_weights(seed=1234, dur_bias=dur_bias, dur_weight_scale=0.05)
    text = make_inference_inputs(seed, t1)
    net = ref_model(w)
    with torch.no_grad():
        mel, ra = net.inference(text)
        # weight-norm folded (bin/inference.py:80) must not change the result
        net.remove_weight_norm()
        mel2, _ = net.inference(text)
    assert torch.equal(mel, mel2)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, t1=t1, dur_bias=dur_bias,
                        mel_pred=mel.numpy(), reconst_alpha=ra.numpy())
    print(name, "T2", mel.shape[1])


def duration_case(name):
    w = orc.make_weights(seed=1234, dur_bias=1.2)
    net = ref_model(w)
    g = torch.Generator().manual_seed(7)
    xs = torch.randn(3, 21, 512, generator=g)
    lens = torch.tensor([21, 13, 5])
    masks = make_pad_mask(lens)
    xs = xs.masked_fill(masks.unsqueeze(-1), 0.0)
    dp = net.duration_predictor
    with torch.no_grad():
        log_d = dp(xs, masks)
        d_round = dp.inference(xs, masks)
        d_float = dp.inference(xs, None, to_round=False)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), xs=xs.numpy(), lens=lens.numpy(),
                        log_d=log_d.numpy(), d_round=d_round.numpy(), d_float=d_float.numpy())
    print(name, d_round[0, :8].tolist())


def length_regulator_cases(name):
    lr = LengthRegulator()
    out = {}
    # docstring example, layers/length_regulator.py:60-73
    x = torch.tensor([[[1.0], [2.0], [3.0]]])
    d = torch.tensor([[1, 2, 3]])
    out["doc_out"] = lr(x, d, torch.tensor([3])).numpy()
    # random ragged batch incl. zero durations, one all-zero row (fix-up :76-78)
    g = torch.Generator().manual_seed(11)
    xs = torch.randn(5, 12, 8, generator=g)
    ds = torch.randint(0, 6, (5, 12), generator=g)
    ilens = torch.tensor([12, 7, 1, 9, 4])
    ds[3, :9] = 0
    out["xs"], out["ds_in"], out["ilens"] = xs.numpy(), ds.numpy().copy(), ilens.numpy()
    ds1 = ds.clone()
    out["out_a1"] = lr(xs, ds1, ilens).numpy()
    out["ds_after_a1"] = ds1.numpy()                 # all-zero row rewritten in place
    ds2 = ds.clone()
    out["out_a13"] = lr(xs, ds2, ilens, alpha=1.3).numpy()
    ds3 = ds.clone()
    out["out_a05"] = lr(xs, ds3, ilens, alpha=0.5).numpy()   # exercises round-half-even
    lr9 = LengthRegulator(pad_value=-9.0)
    out["out_pad9"] = lr9(xs, ds.clone(), ilens).numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, out["out_a1"].shape, out["out_a13"].shape, out["out_a05"].shape)


def mask_cases(name):
    # docstring examples, utils/nets_utils.py:71-76,183-188
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        pad=make_pad_mask([5, 3, 2]).numpy(),
                        non_pad=make_non_pad_mask([5, 3, 2]).numpy())


if __name__ == "__main__":
    torch.set_num_threads(8)
    forward_case("fwd_small", seed=3, t1=[24, 17, 9], t2=[150, 101, 60])
    forward_case("fwd_pad_heavy", seed=4, t1=[40, 6], t2=[260, 37])
    inference_case("inf_small", seed=5, t1=12, dur_bias=float(np.log(6.0)))
    duration_case("dur_small")
    length_regulator_cases("lr_cases")
    mask_cases("masks")
