#!/usr/bin/env python3
"""Micro-benchmark of efts_tap_gemm (fp32 out only) for epilogue-cost experiments."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import efts_oracle as orc
import efficient_tts_b200 as E
from efficient_tts_b200 import workloads as wl

dev = torch.device("cuda", 0)
m = E.EfficientTTSCNN(**wl.MODEL_KWARGS); m.load_state_dict(orc.make_weights(seed=1234)); m = m.eval().to(dev)
eng = m._get_engine()
for kv in os.environ.get("EFTS_OPTS", "").split(","):
    if kv: eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
B, T = 256, 1200
for (K, N, nt) in [(80, 512, 1), (512, 512, 1), (512, 80, 1), (512, 512, 5), (200, 512, 1)]:
    x = torch.randn(B, T, K, device=dev); w = torch.randn(nt, N, K, device=dev) * 0.05
    for _ in range(2): eng.tap_gemm(x, w, ntaps=nt, pad=(nt - 1) // 2)
    torch.cuda.synchronize()
    eng.profile_enable(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # time only the GEMM: tap_gemm = 2 split kernels + gemm; measure whole and split separately
    n = 5
    e0.record()
    for _ in range(n): eng.tap_gemm(x, w, ntaps=nt, pad=(nt - 1) // 2)
    e1.record(); torch.cuda.synchronize()
    tot = e0.elapsed_time(e1) / n
    # split cost estimate: bytes moved
    print("MICRO", json.dumps(dict(K=K, N=N, ntaps=nt, ms_total=round(tot, 4),
          out_MB=B * T * N * 4 / 1e6, in_MB=B * T * K * 4 / 1e6)), flush=True)
