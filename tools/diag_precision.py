#!/usr/bin/env python3
"""Precision diagnostics on the GPU box: (1) sign of the tap-GEMM error (accumulator rounding mode),
(2) full-size C3 / C5 forward against the CPU oracle.  Writes gpurun_out/precision.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import efts_oracle as orc  # noqa: E402
import efficient_tts_b200 as E  # noqa: E402
from efficient_tts_b200 import workloads as wl  # noqa: E402


def main():
    which = sys.argv[1:] or ["bias", "C3", "C5"]
    dev = torch.device("cuda", 0)
    w = orc.make_weights(seed=1234)
    m = E.EfficientTTSCNN(**wl.MODEL_KWARGS)
    m.load_state_dict(w)
    m = m.eval().to(dev)
    eng = m._get_engine()
    for kv in os.environ.get("EFTS_OPTS", "").split(","):
        if kv:
            eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    res = {}
    if "bias" in which:
        g = torch.Generator().manual_seed(5)
        B, T, K, N, nt = 4, 512, 512, 512, 5
        x = torch.randn(B, T, K, generator=g)
        wt = torch.randn(nt, N, K, generator=g) / np.sqrt(K * nt)
        out = eng.tap_gemm(x.to(dev), wt.to(dev), ntaps=nt, pad=2).cpu().double()
        xp = torch.nn.functional.pad(x.double(), (0, 0, 2, 2))
        ref = sum(xp[:, j:j + T] @ wt[j].double().T for j in range(nt))
        f32 = torch.nn.functional.conv1d(x.transpose(1, 2), wt.permute(1, 2, 0).contiguous(), padding=2).transpose(1, 2).double()
        err = out - ref
        e32 = f32 - ref
        big = ref.abs() > 1.0
        res["bias"] = dict(max_err=float(err.abs().max()), rms_err=float(err.pow(2).mean().sqrt()),
                           signed_mean=float((err * ref.sign()).mean()),
                           signed_mean_big=float((err * ref.sign())[big].mean()),
                           rel_shrink_big=float(((err / ref)[big]).mean()),
                           torch_f32_max_err=float(e32.abs().max()), torch_f32_rms=float(e32.pow(2).mean().sqrt()),
                           torch_f32_signed=float((e32 * ref.sign()).mean()))
        print("BIAS", json.dumps(res["bias"]), flush=True)
    for name in ("C2", "C3", "C5"):
        if name not in which:
            continue
        t1, t2 = wl.config_lengths(name)
        text, tl, speech, sl = wl.make_forward_inputs(0, t1, t2)
        out = m(text=text.to(dev), text_lengths=tl.to(dev), speech=speech.to(dev), speech_lengths=sl.to(dev))
        t0 = time.time()
        with torch.no_grad():
            ref = orc.forward(w, text, tl, speech, sl)
        cpu_s = time.time() - t0
        d = {}
        for key, a, b in (("imv", out[2], ref[2]), ("reconst_alpha", out[3], ref[3]), ("mel", out[4], ref[4])):
            diff = (a.cpu() - b).abs()
            d[key] = dict(max=float(diff.max()), p9999=float(diff.flatten().kthvalue(int(diff.numel() * 0.9999)).values),
                          rms=float(diff.pow(2).mean().sqrt()), n_over_1e4=int((diff > 1e-4).sum()), numel=diff.numel())
        if "fp64" in which and name == "C5":
            # the reference's own conditioning: its fp64 restatement vs its fp32 run, and ours vs fp64
            w64 = {k: v.double() for k, v in w.items()}
            _f = torch.Tensor.float
            torch.Tensor.float = lambda self, *a, **k: self.double()
            try:
                with torch.no_grad():
                    r64 = orc.forward(w64, text, tl, speech.double(), sl)
            finally:
                torch.Tensor.float = _f
            for key, i in (("imv", 2), ("reconst_alpha", 3), ("mel", 4)):
                d[key]["ours_vs_fp64_max"] = float((out[i].cpu().double() - r64[i]).abs().max())
                d[key]["ref32_vs_fp64_max"] = float((ref[i].double() - r64[i]).abs().max())
                d[key]["ours_vs_fp64_rms"] = float((out[i].cpu().double() - r64[i]).pow(2).mean().sqrt())
                d[key]["ref32_vs_fp64_rms"] = float((ref[i].double() - r64[i]).pow(2).mean().sqrt())
        d["loss"] = [out[1]["loss"], ref[1]["loss"]]
        d["cpu_oracle_seconds"] = cpu_s
        d["cpu_threads"] = torch.get_num_threads()
        res[name] = d
        print(name, json.dumps(d), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "precision.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
