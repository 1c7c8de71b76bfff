#!/usr/bin/env python3
"""Benchmark of the EFTS-CNN forward path on B200 (contract: the task's bench.py section).

  python bench.py --gpus 1 --steps K --warmup W                  # our CUDA path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                           # the reference's CPU path (oracle port)

Metric (BASELINE.json): mel frames/s of ``EfficientTTSCNN.forward`` at batch=256 mixed-length
(config C3: 256 utterances, 50-200 tokens, 6 frames per token, padded to (200, 1200); 196 950 valid
frames per 256-utterance draw).  One step = one forward over one such batch per GPU.  At N > 1 rank r
runs the C3 draw of seed r (C4 = 8 x C3, weak scaling) through ``DataParallelForward``: the step's only
exchange, the NCCL all-reduce of the loss partial sums and error bits, is inside the timed region.

Every number that comes from a committed profile (ncu traffic, tensor-pipe activity) is printed only when the
profile was taken from the binary that is running: ``profiles/roofline_traffic.json`` carries the hash of the
sources it was captured from and the kernel's template instantiation; both are checked against the library.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from efficient_tts_b200 import workloads as wl  # noqa: E402

WORKLOAD = "C3: batch=256 mixed-length 50-200 tokens, 6 frames/token, 80-bin mel, padded (200,1200)"
UNIT = "mel_frames/s"
METRIC = "mel_frames_per_sec_forward_b256"
TAGS = ["text_conv", "mel_conv", "dec_conv", "linear", "energy_gemm", "softmax_expect", "imv_scan",
        "aligned_pos", "reconstruct", "expand_gemm", "duration", "loss", "embed_split"]

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
# communicator), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved stdout.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tf=float(p["bf16_tflops_sustained"]), tf_burst=float(p["bf16_tflops"]),
                    hbm=float(p["hbm_gbs"]), src="measured (MEASURED_PEAKS.json, sustained bf16)")
    except Exception:
        return dict(tf=1400.0, tf_burst=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.query_ms = []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                t0 = time.perf_counter()
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.query_ms.append((time.perf_counter() - t0) * 1e3)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        return dict(sm_mhz=float(np.median(self.samples)) if self.samples else None,
                    sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.samples),
                    nvml_ms_per_query=float(np.mean(self.query_ms)) if self.query_ms else None)


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def computed_rows(lengths, T, halo):
    """Rows of the 128-row tiles the conv kernel actually computes for one layer (tiles starting at
    or beyond L_b + halo are skipped)."""
    n = 0
    for L in lengths:
        lim = min(T, L + halo)
        n += min(-(-T // 128), -(-lim // 128)) * 128
    return n


def round8(x):
    return (x + 7) // 8 * 8


def cuda_timed(fn, n, dev):
    """Milliseconds per call of ``fn`` over ``n`` back-to-back calls, CUDA events on the current stream."""
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / n


def library_source_sha(eng):
    v = eng.lib.efts_version().decode()
    return v.split(" src ")[-1] if " src " in v else None


def committed_profile(eng, tag):
    """The committed ncu capture of the kernel this library launches under `tag`, or (None, why)."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(path):
        return None, "profiles/roofline_traffic.json is missing"
    try:
        with open(path) as f:
            tj = json.load(f)
    except Exception as exc:
        return None, "profiles/roofline_traffic.json unreadable: %s" % exc
    sha = library_source_sha(eng)
    if tj.get("source_sha16") != sha:
        return None, ("profiles/roofline_traffic.json was captured from sources %s, this library is built from %s: "
                      "its ncu numbers are not quoted" % (tj.get("source_sha16"), sha))
    name = eng.profile_kernel_name(tag) if isinstance(tag, int) else tag
    for k in tj.get("kernels", []):
        if k.get("kernel") == name:
            return k, None
    return None, "no capture of %s in profiles/roofline_traffic.json" % name


# ------------------------------------------------------------------------------------------------
def cpu_reference_forward(state_dict, config, sample_b, seed, steps, warmup, threads):
    """The reference's CPU forward (oracle port: same torch CPU ops, same order) on a draw of the named config.
    Returns (frames/s, seconds per step, sample description)."""
    from oracle import efts_oracle as orc
    torch.set_num_threads(threads)
    t1, t2 = wl.config_lengths(config, seed=seed, batch=sample_b)
    text, tl, speech, sl = wl.make_forward_inputs(seed, t1, t2)
    w = {k: v.detach().cpu() for k, v in state_dict.items()}
    frames = int(sl.sum())
    with torch.no_grad():
        for _ in range(warmup):
            orc.forward(w, text, tl, speech, sl)
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.forward(w, text, tl, speech, sl)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    full = sample_b is None or sample_b == len(wl.config_lengths(config)[0])
    desc = "%s %s: %d utterances, padded (%d,%d), %d valid frames, %d step(s)" % (
        "the full" if full else "a bounded draw of", config, len(t1), text.shape[1], speech.shape[1], frames, steps)
    return frames / dt, dt, desc


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (the oracle port -- the reference is Python on torch CPU
    ops, there is nothing to compile) on the FULL C3 batch, all host threads.  One process: under torchrun rank 0
    runs it alone, so at N > 1 the value is still one host's throughput."""
    if rank != 0:
        return 0
    import efficient_tts_b200 as E
    torch.manual_seed(1234)
    sd = E.EfficientTTSCNN(**wl.MODEL_KWARGS).state_dict()
    threads = os.cpu_count() or 1
    sample = None if args.cpu_sample in (0, 256) else args.cpu_sample
    fps, dt, desc = cpu_reference_forward(sd, "C3", sample, 0, max(1, args.steps), max(1, min(args.warmup, 1)), threads)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": desc, "utterances_per_step": sample or 256,
                       "note": "one CPU process on rank 0 whatever N is (the host does not scale with the GPU count)"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
def time_forward_config(eng, dev, name, steps=5, warmup=2):
    """Device-resident forward time of another BASELINE config (C2 / C5): ms per step and valid frames/s."""
    t1, t2 = wl.config_lengths(name)
    text, tl, speech, sl = (t.to(dev) for t in wl.make_forward_inputs(0, t1, t2))
    for _ in range(warmup):
        eng.forward(text, tl, speech, sl)
    ms = cuda_timed(lambda: eng.forward(text, tl, speech, sl), steps, dev)
    frames = int(sum(t2))
    flops = forward_flops(len(t1), max(t1), max(t2))
    return {"config": "%s: B=%d, T1=%d, T2=%d (padded)" % (name, len(t1), max(t1), max(t2)), "ms_per_step": ms,
            "valid_frames_per_s": frames / (ms * 1e-3), "padded_algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
            "steps": steps}


def forward_flops(B, T1, T2, C=512, odim=80):
    """Single-pass algorithmic FLOPs of one padded forward."""
    conv = lambda T, k, n: 2.0 * B * T * C * C * k * n
    return (conv(T1, 5, 5) + conv(T2, 5, 3) + conv(T2, 5, 6) + conv(T1, 3, 2) + 2 * 2.0 * B * T1 * C * C +
            2.0 * B * T2 * odim * C * 2 + 2 * 2.0 * B * T1 * T2 * C)


def time_length_regulator(dev, peaks):
    """LengthRegulator.forward at the C3 shape (layers/length_regulator.py:35-79): bytes per SURVEY.md 8d."""
    from efficient_tts_b200 import engine as E
    t1, _ = wl.config_lengths("C3", seed=0)
    B, T1, D = len(t1), max(t1), 512
    g = torch.Generator().manual_seed(3)
    xs = torch.randn(B, T1, D, generator=g).to(dev)
    ds = torch.randint(3, 10, (B, T1), generator=g)
    il = torch.tensor(t1, dtype=torch.int64)
    for b in range(B):
        ds[b, t1[b]:] = 0
    ds, il = ds.to(dev), il.to(dev)
    for _ in range(3):
        out, idx = E.length_regulator(xs, ds, il, return_index=True)
    ms = cuda_timed(lambda: E.length_regulator(xs, ds, il, return_index=True), 10, dev)
    tout = int(ds.sum())
    by = 4 * D * (int(il.sum()) + tout) + 8 * B * T1 + 8 * tout
    padded = by - 4 * D * tout + 4 * D * B * out.shape[1]
    return {"config": "LengthRegulator B=%d T1=%d D=%d -> %d frames (Tout_max %d)" % (B, T1, D, tout, out.shape[1]),
            "ms_per_call": ms, "algorithmic_bytes": by, "bytes_incl_padded_rows": padded,
            "achieved_gbs": by / ms / 1e6, "achieved_gbs_incl_padded_rows": padded / ms / 1e6, "peak_gbs": peaks["hbm"],
            "frac": by / ms / 1e6 / peaks["hbm"], "frac_incl_padded_rows": padded / ms / 1e6 / peaks["hbm"],
            "note": "the call includes the plan kernel, the Tout read-back and the output allocation (the reference's "
                    "contract: the padded length is data dependent); bytes per SURVEY.md 8d, rows up to max_b Tout_b are "
                    "written as padding (pad_list)"}


def time_frontend(dev, peaks):
    """Log-mel front-end (datasets/meldataset.py:49-82) on the audio of a C3 batch: 256 utterances
    of 300-1200 frames (76 800 - 307 200 samples), one batched call."""
    from efficient_tts_b200.frontend import LogMelFrontend
    t1, t2 = wl.config_lengths("C3", seed=0)
    lengths = [f * 256 for f in t2]
    g = torch.Generator().manual_seed(9)
    audio = (torch.rand(len(t2), max(lengths), generator=g) * 2 - 1) * 0.5
    lens = torch.tensor(lengths)
    audio = (audio * (torch.arange(max(lengths))[None] < lens[:, None])).to(dev)
    lens = lens.to(dev)
    fe = LogMelFrontend(dev)
    for _ in range(3):
        fe(audio, lens)
    n0 = fe.launch_count()
    ms = cuda_timed(lambda: fe(audio, lens, check=False), 10, dev)
    launches = (fe.launch_count() - n0) // 10
    frames = int(sum(t2))
    rows = len(t2) * max(t2)                             # frames of the padded batch the STFT GEMM computes
    kept = 372                                           # bins some mel filter weighs (fmax 8 kHz of 11.025): the GEMM skips the others
    flops_ref = 2.0 * rows * 1024 * 1024 + 2.0 * rows * 520 * 80     # the reference's full transform + projection
    flops = 2.0 * rows * (2 * kept) * 1024 + 2.0 * rows * kept * 80  # what this path multiplies
    by = 4.0 * sum(lengths) + 4.0 * 80 * frames         # audio in, log-mel out
    return {"config": "mel_spectrogram (n_fft 1024, hop 256, 80 mels) on a C3 batch: %d utterances, %d valid frames, "
                      "%.1f M samples" % (len(t2), frames, sum(lengths) / 1e6),
            "ms_per_call": ms, "valid_frames_per_s": frames / (ms * 1e-3), "gpu_launches_per_call": launches,
            "algorithmic_tflops": flops / (ms * 1e-3) / 1e12, "frac_of_tensor_peak": flops / (ms * 1e-3) / 1e12 / peaks["tf"],
            "executed_frac": 3 * flops / (ms * 1e-3) / 1e12 / peaks["tf"],
            "reference_transform_tflops_equivalent": flops_ref / (ms * 1e-3) / 1e12,
            "io_bytes": by, "io_gbs": by / ms / 1e6,
            "note": "STFT as a 4-tap GEMM over 256-sample chunks (K = 4 x 256; N = 744 = the (re, im) column pairs of the 372 "
                    "bins the mel filter bank weighs -- the other 141 are multiplied by zero in the reference) with the "
                    "magnitude in its epilogue, and the mel projection (K = 372, N = 80, log in the epilogue) on the tcgen05 "
                    "tap-GEMM, split-fp16 x 3 passes; FLOPs = what this path multiplies on the padded batch"}

def time_training_slice(dev, peaks):
    """Training slice: forward-with-saved-activations and backward of the 6-layer decoder stack at the
    C3 padded shape (B = 256, T = 1200, 512 channels, k = 5), raw library calls."""
    from efficient_tts_b200.engine import train_context
    tc = train_context(dev)
    g = torch.Generator().manual_seed(12)
    L, B, T, C, k = 6, 256, 1200, 512, 5
    x = torch.randn(B, T, C, generator=g).to(dev)
    w = (torch.randn(L, C, C, k, generator=g) / np.sqrt(C * k)).to(dev)
    b = (torch.randn(L, C, generator=g) * 0.1).to(dev)
    grad = torch.randn(B, T, C, generator=g).to(dev)
    # each call allocates its 8 GB of saved activations through torch's caching allocator, which now and then has to go
    # to the driver (tens of ms): the median of five single-call timings is reported
    acts, us = tc.resconv_fwd(x, w, b)
    del acts, us
    ms_f = float(np.median([cuda_timed(lambda: tc.resconv_fwd(x, w, b), 1, dev) for _ in range(5)]))
    acts, us = tc.resconv_fwd(x, w, b)
    tc.resconv_bwd(grad, acts, us, w)
    n0 = tc.launch_count()
    ms_b = float(np.median([cuda_timed(lambda: tc.resconv_bwd(grad, acts, us, w), 1, dev) for _ in range(5)]))
    launches = (tc.launch_count() - n0) // 5
    fl = 2.0 * B * T * C * C * k * L                      # one conv pass over the stack (single-pass algorithmic)
    # second slice: the duration predictor at the C3 text shape (B = 256, T1 = 200, 2 layers, k = 3), raw library calls
    dL, dT, dk = 2, 200, 3
    dx = torch.randn(B, dT, C, generator=g).to(dev)
    dw = (torch.randn(dL, C, C, dk, generator=g) / np.sqrt(C * dk)).to(dev)
    dvec = lambda *s: (torch.randn(*s, generator=g) * 0.1).to(dev)
    dcb, dlg, dlb, dhw, dhb = dvec(dL, C), dvec(dL, C) + 1.0, dvec(dL, C), dvec(C), dvec(1)
    dgrad = torch.randn(B, dT, generator=g).to(dev)
    _, dacts, dus = tc.duration_fwd(dx, dw, dcb, dlg, dlb, dhw, dhb)
    dp_f = cuda_timed(lambda: tc.duration_fwd(dx, dw, dcb, dlg, dlb, dhw, dhb), 3, dev)
    tc.duration_bwd(dgrad, dacts, dus, dw, dlg, dhw)
    dp_b = cuda_timed(lambda: tc.duration_bwd(dgrad, dacts, dus, dw, dlg, dhw), 3, dev)
    duration = {"config": "DurationPredictor (2 x [Conv1d k=3, ReLU, LayerNorm], Linear 512 -> 1), B=%d, T1=%d" % (B, dT),
                "forward_ms": dp_f, "backward_ms": dp_b}
    return {"duration_predictor": duration, "config": "ResConvBlock x %d layers, B=%d, T=%d (padded C3 decoder shape), fp32 weights passed per call" % (L, B, T),
            "forward_ms": ms_f, "backward_ms": ms_b, "backward_launches": launches,
            "forward_tflops": fl / (ms_f * 1e-3) / 1e12, "backward_tflops": 2 * fl / (ms_b * 1e-3) / 1e12,
            "backward_frac_of_tensor_peak": 2 * fl / (ms_b * 1e-3) / 1e12 / peaks["tf"],
            "note": "backward = data gradient (one tap-GEMM per layer) + weight gradient (k position-reduction GEMMs per "
                    "layer, the layer input read in place as an MN-major operand, the tap as a row offset) + bias "
                    "gradient; algorithmic FLOPs 2 x forward, single pass (3 passes executed)"}


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    import efficient_tts_b200 as E
    from efficient_tts_b200.data_parallel import DataParallelForward, combine_loss_partials
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py measures the CUDA path; no CUDA device is visible (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)                      # same default init on every rank == replicated weights
    model = E.EfficientTTSCNN(**wl.MODEL_KWARGS).eval()
    state = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(dev)
    eng = model._get_engine()
    for kv in os.environ.get("EFTS_BENCH_OPTS", "").split(","):     # A/B switches for experiments
        if kv:
            eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))

    t1, t2 = wl.config_lengths("C3", seed=rank)
    host = [t.pin_memory() for t in wl.make_forward_inputs(rank, t1, t2)]
    text, tl, speech, sl = (t.to(dev) for t in host)
    frames = int(host[3].sum())
    B, T1p, T2p = text.shape[0], text.shape[1], speech.shape[1]
    dp = DataParallelForward(model.forward_shard) if world > 1 else None

    def step(a=text, b=tl, c=speech, d=sl):
        """One step, nothing read back: the forward over this rank's 256 utterances and, at N > 1, the all-reduce
        of the loss partial sums / error bits that turns the shards into the B = 256 N forward of C4."""
        if dp is None:
            return eng.forward(a, b, c, d)
        return dp.launch_local(a, b, c, d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput (value) ---------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # one untimed profiled pass creates every CUDA event the per-kernel breakdown needs; the timed pass reuses them
    eng.profile_enable(0x1FFF)
    for _ in range(args.steps):
        step()
    barrier()
    t_host0 = time.perf_counter()                # host cost of enqueueing one step (launch queue empty: 2 steps fit)
    for _ in range(2):
        step()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / 2
    barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    eng.profile_enable(0x1FFF)
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    barrier()
    ms_sampled = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    prof = {tag: eng.profile_read(tag) for tag in range(13)}
    dec_kernel = eng.profile_kernel_name(2)
    eng.profile_enable(0)
    clocks = sampler.stop()
    # The same K steps once more, immediately, without the NVML sampler thread: on some boxes every NVML query
    # takes ~10 ms and stalls the GPU's work submission (seen as a timed pass 30-70 % slower than the sum of its own
    # kernel times while the SM clock reads idle-high).  Both timings are reported; `value` uses the sampled pass
    # unless the queries were slow AND the unsampled pass is more than 3 % faster.
    barrier()
    ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev4.record()
    for _ in range(args.steps):
        out = step()
    ev5.record()
    barrier()
    ms_unsampled = ev4.elapsed_time(ev5)
    slow_nvml = (clocks.get("nvml_ms_per_query") or 0.0) > 1.0
    use_unsampled = slow_nvml and ms_unsampled < 0.97 * ms_sampled
    ms = ms_unsampled if use_unsampled else ms_sampled
    timing = {"ms_per_step_clock_sampled_pass": ms_sampled / args.steps,
              "ms_per_step_unsampled_pass": ms_unsampled / args.steps,
              "reported": "unsampled pass (NVML queries took %.1f ms each and stalled the sampled pass; clocks are "
                          "from the sampled pass of the same K steps run immediately before)" % clocks["nvml_ms_per_query"]
              if use_unsampled else "clock-sampled pass"}
    dp_check = None
    if dp is not None:
        # the reduced loss of the last timed step, and -- on rank 0 -- the same B = 256 N batch run shard by shard in
        # ONE process: the two must agree (the all-reduce really produced the loss of the whole C4 batch)
        loss_dp, stats_dp = dp.finish(out[3])
        if rank == 0:
            acc = np.zeros(4, dtype=np.float64)
            for r in range(world):
                a1, a2 = wl.config_lengths("C3", seed=r)
                sh = [t.to(dev) for t in wl.make_forward_inputs(r, a1, a2)]
                acc += eng.forward(*sh)[3][3:7].double().cpu().numpy()
            loss_1p = combine_loss_partials(acc)[0]
            dp_check = {"loss_all_reduced": loss_dp, "loss_single_process_B%d" % (B * world): loss_1p,
                        "rel_diff": abs(loss_dp - loss_1p) / abs(loss_1p)}
            assert dp_check["rel_diff"] <= 1e-5, dp_check

    # ---- end to end through the public API: pinned host inputs -> H2D -> forward -> stats read-back.
    # Like a prefetching loader (the reference trains with pin_memory + non_blocking copies), the copy of
    # step i + 1 is issued on a second stream while step i computes; every step's inputs cross PCIe inside
    # the timed region and every step ends with the host reading its loss statistics.  A second variant also brings
    # mel_pred back to pinned host memory every step (what the reference's _eval_step plots, trainers/...:218).
    copy_stream = torch.cuda.Stream(dev)
    dbuf = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]
    mel_host = [torch.empty(B, T2p, wl.ODIM, dtype=torch.float32).pin_memory() for _ in range(2)]

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            for d, h in zip(dbuf[slot], host):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_run(n, mel_back):
        stats = None
        ready = issue_copy(0)
        pending = None
        for i in range(n):
            torch.cuda.current_stream(dev).wait_event(ready)
            d = dbuf[i % 2]
            if i + 1 < n:
                ready = issue_copy((i + 1) % 2)      # slot (i+1)%2 was released by step i-1's read-back
            if dp is None:
                loss, stats, imv, ra, mel, _ = model(text=d[0], text_lengths=d[1], speech=d[2], speech_lengths=d[3])
            else:
                imv, ra, mel, part = dp.launch_local(d[0], d[1], d[2], d[3])
                loss, stats = dp.finish(part)
            if mel_back:
                # D2H of this step's mel_pred on the copy stream, overlapping the next step's compute
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(dev))
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    mel_host[i % 2].copy_(mel, non_blocking=True)
                    mel.record_stream(copy_stream)
                    fin = torch.cuda.Event()
                    fin.record(copy_stream)
                if pending is not None:
                    pending.synchronize()            # the previous step's mel is on the host
                pending = fin
        if pending is not None:
            pending.synchronize()
        return stats
    e2e_ms = {}
    for mel_back in (False, True):
        e2e_run(2, mel_back)
        barrier()
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2.record()
        stats = e2e_run(args.steps, mel_back)
        ev3.record()
        barrier()
        e2e_ms[mel_back] = ev2.elapsed_time(ev3)
    h2d = sum(t.numel() * t.element_size() for t in host)
    mel_bytes = B * T2p * wl.ODIM * 4

    # ---- max over ranks, totals -----------------------------------------------------------------
    t = torch.tensor([ms, e2e_ms[False], e2e_ms[True]], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(frames)], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        mine = torch.tensor([ms / args.steps, clocks.get("sm_mhz") or 0.0], dtype=torch.float64, device=dev)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        per_rank = {"ms_per_step": [float(g[0]) for g in gathered], "sm_mhz": [float(g[1]) for g in gathered]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_e2e, ms_e2e_mel = (float(v) for v in t.tolist())
    total_frames = float(tot.item())
    value = total_frames * args.steps / (ms * 1e-3)
    e2e_value = total_frames * args.steps / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        peaks = load_peaks()
        # dominant kernel: the decoder Conv1d layer (tag 2).  Algorithmic FLOPs per launch =
        # 2 * rows * 512 * 512 * 5, rows = rows of the tiles the launch computes.
        dec_ms, dec_n = prof[2]
        n_dec = 6
        rows = np.mean([computed_rows(t2, T2p, 2 * (n_dec - 1 - l)) for l in range(n_dec)])
        flops_per_launch = 2.0 * rows * 512 * 512 * 5
        ach = flops_per_launch / (dec_ms / max(dec_n, 1) * 1e-3) / 1e12 if dec_n else None
        cap, why = committed_profile(eng, dec_kernel)
        roof = {"bound": "tensor",
                "kernel": "%s (CTA-pair tcgen05 tap-GEMM, fused-B) on the decoder Conv1d layers (k=5, 512->512)" % dec_kernel,
                "achieved": ach, "peak": peaks["tf"], "unit": "TFLOP/s", "frac": (ach / peaks["tf"]) if ach else None,
                "traffic": cap.get("dram_bytes") if cap else None,
                "tensor_pipe_active_pct_ncu": cap.get("tensor_pipe_active_pct") if cap else None,
                "ncu_duration_ms": cap.get("duration_ms") if cap else None,
                "ncu_source": cap.get("source") if cap else why,
                "library_source_sha16": library_source_sha(eng),
                "peak_source": peaks["src"], "passes": 3,
                "executed_frac": (3 * ach / peaks["tf"]) if ach else None,
                "flops_per_launch": flops_per_launch, "launches_timed": dec_n,
                "avg_launch_ms": dec_ms / max(dec_n, 1),
                "share_of_step": dec_ms / ms if ms else None,
                "algorithmic_bytes_per_launch": 8.0 * rows * 512,
                "note": "achieved = single-pass algorithmic FLOPs; the split-fp16 scheme executes 3 tensor passes "
                        "(executed_frac = 3 x frac, against the power-capped cuBLAS bf16 rate); algorithmic bytes = "
                        "read 4 B + write 4 B per element of the computed rows"}
        breakdown = {TAGS[k]: round(v[0] / args.steps, 4) for k, v in prof.items()}
        # IMV (HBM-bound) kernels: algorithmic bytes per forward against the measured HBM peak, two definitions.
        m2, m1r = B * T2p, B * T1p
        live2 = float(sum(t2))
        n_part = -(-round8(T1p) // 128)
        scan_bytes = 16 * n_part * live2 + 4 * m2 + 8 * m2
        aligned_bytes = 4 * live2 + 4 * m1r
        recon_bytes = 4 * m1r + 4 * B * T1p * T2p + 4 * live2 * round8(T1p)
        imv_bytes = scan_bytes + aligned_bytes + recon_bytes
        imv_ms = sum(prof[k][0] for k in (5, 6, 7, 8)) / args.steps
        rec_ms = prof[8][0] / args.steps
        rcap, rwhy = committed_profile(eng, "reconstruct_alignment_rows_kernel")
        # fused minimum of the whole block = 4 [B T2 D (mel_h) + 2 B T1 D (key, value) + B T2 (imv)
        # + B T1 (e) + B T1 T2 (reconst_alpha) + B T2 D (expanded)], padded sizes, over energy + scan + aligned
        # positions + reconstruction + expansion
        block_bytes = 4.0 * (m2 * 512 + 2 * m1r * 512 + m2 + m1r + B * T1p * T2p + m2 * 512)
        block_ms = sum(prof[k][0] for k in (4, 5, 6, 7, 8, 9)) / args.steps
        hbm = {"definition": "per-kernel bytes actually required by the three streaming kernels (partials in, imv / e / "
                             "reconst_alpha + operand planes out)",
               "kernels": "imv_scan_block + aligned_positions_block + reconstruct_alignment_rows", "bytes_per_step": imv_bytes,
               "ms_per_step": imv_ms, "achieved_gbs": imv_bytes / (imv_ms * 1e-3) / 1e9 if imv_ms else None,
               "peak_gbs": peaks["hbm"], "frac": (imv_bytes / (imv_ms * 1e-3) / 1e9 / peaks["hbm"]) if imv_ms else None,
               "note": "scan and aligned positions move ~11 MB and are latency / exp-throughput bound; the HBM-bound "
                       "kernel of the chain is the Gaussian reconstruction (97 % of the bytes)",
               "reconstruct": {"bytes": recon_bytes, "ms": rec_ms, "traffic": rcap.get("dram_bytes") if rcap else None,
                               "ncu_source": rcap.get("source") if rcap else rwhy,
                               "achieved_gbs": recon_bytes / (rec_ms * 1e-3) / 1e9 if rec_ms else None,
                               "frac": (recon_bytes / (rec_ms * 1e-3) / 1e9 / peaks["hbm"]) if rec_ms else None,
                               "frac_of_8tbs_nominal": (recon_bytes / (rec_ms * 1e-3) / 1e9 / 8000.0) if rec_ms else None},
               "block_survey_8d": {"definition": "SURVEY.md 8d fused-minimum bytes of the whole alignment block (padded sizes) "
                                                 "over energy GEMM + scan + aligned positions + reconstruction + expansion GEMM",
                                   "bytes": block_bytes, "ms": block_ms,
                                   "achieved_gbs": block_bytes / (block_ms * 1e-3) / 1e9 if block_ms else None,
                                   "frac": (block_bytes / (block_ms * 1e-3) / 1e9 / peaks["hbm"]) if block_ms else None,
                                   "frac_of_8tbs_nominal": (block_bytes / (block_ms * 1e-3) / 1e9 / 8000.0) if block_ms else None}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "dtype_note": "fp32 in/out; split-fp16 (hi/lo) tensor-core operands, fp32 accumulation",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "utterances_per_gpu": B, "valid_frames_per_gpu": frames,
                           "padded": [T1p, T2p], "l2": "working set per step ~3.3 GB >> 126 MB L2 (no flush needed)",
                           "parallelism": ("dp%d: utterance shards (C3 draw of seed r on rank r = C4), one NCCL all-reduce of "
                                           "the loss partial sums + error bits per step inside the timed region" % world)
                           if world > 1 else "dp1 (single GPU, no collective)"},
                "padded_frames_per_s": float(B * T2p) * world * args.steps / (ms * 1e-3),
                "clocks": clocks, "timing": timing, "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * 4,
                        "ms_per_step": ms_e2e / args.steps, "stats": stats,
                        "outputs": "loss statistics read back every step (what the reference's _train_step consumes); "
                                   "imv / reconst_alpha / mel_pred stay device-resident",
                        "pipeline": "H2D of step i+1 on a copy stream overlaps step i; loss read back every step",
                        "with_mel_pred_d2h": {"value": total_frames * args.steps / (ms_e2e_mel * 1e-3), "unit": UNIT,
                                              "ms_per_step": ms_e2e_mel / args.steps,
                                              "d2h_bytes_per_step": 8 * 4 + mel_bytes,
                                              "note": "additionally copies mel_pred to pinned host memory every step on the "
                                                      "copy stream (what _eval_step plots, trainers/efficient_tts_trainer.py:218)"}},
                "roofline": roof, "roofline_hbm_imv": hbm, "kernel_ms_per_step": breakdown}
        if per_rank is not None:
            line["per_rank"] = per_rank
            line["data_parallel_check"] = dp_check
    # ---- extras on rank 0 at N = 1: other configs, RTF at batch 1 (C1), vocoder, CPU baselines -------------------
    if rank == 0 and world == 1 and args.no_extras:
        line["cpu_baseline"] = None
    elif rank == 0 and world == 1:
        try:
            line["other_configs"] = {"C2": time_forward_config(eng, dev, "C2"), "C5": time_forward_config(eng, dev, "C5")}
            line["length_regulator"] = time_length_regulator(dev, load_peaks())
            line["frontend"] = time_frontend(dev, load_peaks())
            line["training_slice"] = time_training_slice(dev, load_peaks())
            torch.cuda.empty_cache()
        except Exception as exc:
            line["other_configs"] = {"error": str(exc)[:200]}
        try:
            extras_c1_and_vocoder(line, E, state, dev, args)
        except Exception as exc:  # the headline line must still print
            line["rtf_batch1"] = {"error": str(exc)[:200]}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = None if args.cpu_sample in (0, 256) else args.cpu_sample
            fps, dt, desc = cpu_reference_forward(state, "C3", sample, 0, 1, 1, threads)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc,
                                    "seconds_per_sample": dt}
            try:
                # the reference's recipes pin OMP_NUM_THREADS=1 (egs/lj/path.sh:13, distributed/launch.py:84-85):
                # the same CPU path on one thread, on an 8-utterance draw; and its B = 1 synthesis latency (C1)
                fps1, dt1, desc1 = cpu_reference_forward(state, "C3", 8, 0, 1, 0, 1)
                line["cpu_baseline"]["one_thread"] = {"value": fps1, "sample": desc1, "seconds_per_sample": dt1}
                from oracle import efts_oracle as orc
                torch.set_num_threads(threads)
                w_c1 = {k: v.detach().cpu() for k, v in wl.c1_weights_patch(state).items()}
                t_c1 = wl.make_inference_inputs(0, 64)
                with torch.no_grad():
                    orc.inference(w_c1, t_c1)
                    t0 = time.perf_counter()
                    for _ in range(3):
                        mel_c, _ = orc.inference(w_c1, t_c1)
                    dt_c = (time.perf_counter() - t0) / 3
                line["cpu_baseline"]["c1_inference"] = {"ms": dt_c * 1e3, "frames": int(mel_c.shape[1]),
                                                        "rtf_mel_only": dt_c / (mel_c.shape[1] * 256 / 22050.0)}
            except Exception as exc:
                line["cpu_baseline"]["extras_error"] = str(exc)[:200]
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def extras_c1_and_vocoder(line, E, state, dev, args):
    """RTF at batch 1 (C1), batched synthesis, the HiFi-GAN generator and text -> waveform (rank 0, N = 1)."""
    def wall(fn, n, warm=3):
        for _ in range(warm):
            r = fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(n):
            r = fn()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / n, r

    mc1 = E.EfficientTTSCNN(**wl.MODEL_KWARGS)
    mc1.load_state_dict(wl.c1_weights_patch(state))
    mc1 = mc1.eval().to(dev)
    eng1 = mc1._get_engine()
    txt = wl.make_inference_inputs(0, 64).to(dev)
    n0 = eng1.launch_count()
    mc1.inference(txt)
    c1_launches = eng1.launch_count() - n0
    dt, (mel, _) = wall(lambda: mc1.inference(txt), 50)
    line["rtf_batch1"] = {"config": "C1: inference, B=1, 64 phonemes -> %d frames" % mel.shape[1],
                          "ms": dt * 1e3, "rtf_mel_only": dt / (mel.shape[1] * 256 / 22050.0),
                          "frames_per_s": mel.shape[1] / dt, "gpu_launches_per_call": c1_launches,
                          "path": "resident layer-stack kernels (stack_sm100.cuh): phase 1, reconstruct, expand, phase 2",
                          "note": "wall time of the public call incl. the T2 read-back and the trailing error-flag read; "
                                  "mel-only RTF; the reference's RTF also includes HiFi-GAN (bin/inference.py:100-111)"}
    eng1.set_option("stack", 0)
    dt_pl, _ = wall(lambda: mc1.inference(txt), 30)
    eng1.set_option("stack", 1)
    line["rtf_batch1"]["ms_one_launch_per_layer"] = dt_pl * 1e3
    # batched variable-length synthesis: 64 utterances of 32-64 tokens in one call
    g = torch.Generator().manual_seed(7)
    lens = torch.randint(32, 65, (64,), generator=g).to(dev)
    btxt = torch.randint(0, wl.NUM_SYMBOLS, (64, 64), generator=g).to(dev)
    dt, (bmel, blen, _) = wall(lambda: mc1.inference_batch(btxt, lens), 10, 2)
    line["inference_batch64"] = {"config": "inference_batch, 64 utterances x 32-64 tokens -> %d frames" % int(blen.sum()),
                                 "ms": dt * 1e3, "frames_per_s": float(blen.sum()) / dt,
                                 "rtf_mel_only": dt / (float(blen.sum()) * 256 / 22050.0)}
    # the step right after the path: HiFi-GAN V1 generator, and the text -> waveform RTF
    # the reference defines (bin/inference.py:100-111 times inference + vocoder)
    from efficient_tts_b200.vocoder import Generator
    voc = Generator(wl.AttrDict(wl.HIFIGAN_V1))
    voc.load_state_dict(wl.vocoder_state_dict())
    voc = voc.eval().to(dev)
    vmel = wl.make_mel(1, 16, 800).to(dev)
    dt, wav = wall(lambda: voc(vmel), 5, 2)
    vline = {"config": "HiFi-GAN V1 generator, 16 x 800 frames -> %d samples" % wav.numel(),
             "ms": dt * 1e3, "samples_per_s": wav.numel() / dt, "rtf": dt / (wav.numel() / 22050.0)}
    vfl = vocoder_flops(wl.HIFIGAN_V1, 16, 800)
    vpk = load_peaks()
    vline["roofline"] = {"bound": "tensor", "achieved": vfl / dt / 1e12, "peak": vpk["tf"], "unit": "TFLOP/s",
                         "frac": vfl / dt / 1e12 / vpk["tf"], "flops_per_call": vfl, "passes": 3,
                         "note": "single-pass algorithmic FLOPs of the 78 convolutions over the whole forward "
                                 "(2 * Cin * Cout * k per output sample; transposed convs 2 * Cin * Cout * k per "
                                 "input sample); the split-fp16 scheme executes 3 passes"}
    dt, wav1 = wall(lambda: voc(mc1.inference(txt)[0].transpose(1, 2)), 20)
    vline["text_to_wave_c1"] = {"config": "inference(64 phonemes) + generator, B=1 -> %d samples" % wav1.shape[-1],
                                "ms": dt * 1e3, "rtf": dt / (wav1.shape[-1] / 22050.0)}
    if not args.no_cpu_baseline:
        from oracle import hifigan_oracle as hor
        torch.set_num_threads(os.cpu_count() or 1)
        cmel = wl.make_mel(1, 16, 800)             # the same 16 x 800 frames the GPU arm runs
        cw = wl.vocoder_state_dict()
        with torch.no_grad():
            hor.generator_forward(cw, cmel[:1, :, :64])
            t0 = time.perf_counter()
            cy = hor.generator_forward(cw, cmel)
            cdt = time.perf_counter() - t0
        vline["cpu_baseline"] = {"value": cy.numel() / cdt, "unit": "samples/s", "cores": os.cpu_count() or 1,
                                 "kind": "port", "sample": "the same 16 x 800 frames (%d samples), 1 pass" % cy.numel(),
                                 "seconds": cdt, "rtf": cdt / (cy.numel() / 22050.0)}
    line["vocoder"] = vline


def vocoder_flops(h, batch, frames):
    """Algorithmic FLOPs of one Generator.forward (vocoders/hifigan_model.py:120-136)."""
    c = h["upsample_initial_channel"]
    L = frames
    fl = 2.0 * h["num_mels"] * c * 7 * L
    for u, k in zip(h["upsample_rates"], h["upsample_kernel_sizes"]):
        fl += 2.0 * c * (c // 2) * k * L              # ConvTranspose1d: every input sample meets all k taps
        c //= 2
        L *= u
        for rk, dil in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
            fl += 2 * len(dil) * 2.0 * c * c * rk * L
    fl += 2.0 * c * 1 * 7 * L
    return fl * batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=256,
                    help="utterances of the C3 draw the CPU legs run (256 = the full batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline C3 measurement (kernel experiments)")
    args = ap.parse_args()
    capture_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world != args.gpus:
        if args.gpus == 1 and world == 1:
            pass
        else:
            raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
    return run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
