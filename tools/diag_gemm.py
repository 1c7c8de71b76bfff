#!/usr/bin/env python3
"""GPU bring-up diagnostics: every tap-GEMM variant in its own process (a trapped kernel poisons
the CUDA context), then a timing of the named bench workload.  Writes gpurun_out/diag.json."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASE = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %(root)r)
from oracle import efts_oracle as orc
import efficient_tts_b200 as E
cfg = %(cfg)r
dev = torch.device("cuda", 0)
m = E.EfficientTTSCNN(num_symbols=76, dropout_rate=0.0, use_masking=True, sigma=0.01)
m.load_state_dict(orc.make_weights(seed=1234)); m = m.eval().to(dev)
eng = m._get_engine()
for k, v in cfg.get("opts", {}).items(): eng.set_option(k, v)
g = torch.Generator().manual_seed(5)
B, T, K, N, nt = cfg["B"], cfg["T"], cfg["K"], cfg["N"], cfg["ntaps"]
x = torch.randn(B, T, K, generator=g); w = torch.randn(nt, N, K, generator=g) / np.sqrt(K * nt)
out = eng.tap_gemm(x.to(dev), w.to(dev), ntaps=nt, pad=(nt - 1) // 2)
torch.cuda.synchronize()
out = out.cpu()
xp = torch.nn.functional.pad(x.double(), (0, 0, (nt - 1) // 2, (nt - 1) // 2))
ref = sum(xp[:, j:j + T] @ w[j].double().T for j in range(nt))
err = (out.double() - ref).abs()
print("RESULT " + json.dumps(dict(cfg=cfg, max_err=float(err.max()), mean_err=float(err.mean()),
      ref_absmax=float(ref.abs().max()), nan=int(torch.isnan(out).sum()))))
'''

V1 = dict(gemm_version=1, amode=0)
V2S = dict(gemm_version=2, pair=0, chunk_kb=1)
V2P = dict(gemm_version=2, pair=1, chunk_kb=1)
SHAPES = [
    dict(B=1, T=128, K=64, N=64, ntaps=1),
    dict(B=1, T=128, K=512, N=256, ntaps=1),
    dict(B=3, T=77, K=80, N=512, ntaps=1),
    dict(B=2, T=300, K=512, N=80, ntaps=1),
    dict(B=2, T=260, K=512, N=512, ntaps=5),
    dict(B=3, T=50, K=512, N=512, ntaps=3),
    dict(B=64, T=1200, K=512, N=512, ntaps=5),
]
CASES = [dict(sh, opts=o) for o in (V2S, V2P) for sh in SHAPES] + [
    dict(B=64, T=1200, K=512, N=512, ntaps=5, opts=dict(gemm_version=2, pair=1, chunk_kb=2)),
    dict(B=64, T=1200, K=512, N=512, ntaps=5, opts=dict(gemm_version=2, pair=1, chunk_kb=0)),
    dict(B=64, T=1200, K=512, N=512, ntaps=5, opts=V1),
]


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = []
    for cfg in CASES:
        code = CASE % dict(root=ROOT, cfg=cfg)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            res = json.loads(line[0][7:]) if line else dict(cfg=cfg, error=(r.stderr or r.stdout)[-1500:])
        except subprocess.TimeoutExpired:
            res = dict(cfg=cfg, error="timeout")
        res["secs"] = round(time.time() - t0, 1)
        print(json.dumps(res), flush=True)
        results.append(res)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
