"""Data-parallel forward on real GPUs (needs >= 2 devices; skipped otherwise): every rank runs its
utterance shard with the global padded dims, the 4-float loss partials are all-reduced over NCCL and
the gathered outputs are bitwise the single-GPU outputs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import efficient_tts_b200 as E
from efficient_tts_b200 import workloads as wl
from efficient_tts_b200.data_parallel import DataParallelForward, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1234)
m = E.EfficientTTSCNN(**wl.MODEL_KWARGS).eval().to(dev)
text, tl, speech, sl = (t.to(dev) for t in wl.make_forward_inputs(61, [40, 17, 33, 8, 25, 12, 40], [230, 90, 200, 40, 150, 70, 240]))
dp = DataParallelForward(m.forward_shard)
loss, stats, imv, ra, mel = dp(text, tl, speech, sl, gather_outputs=True)
full = m(text=text, text_lengths=tl, speech=speech, speech_lengths=sl)
assert torch.equal(imv, full[2]) and torch.equal(ra, full[3]) and torch.equal(mel, full[4]), "gathered outputs differ"
for k in ("loss", "mel_loss", "duration_loss"):
    assert abs(stats[k] - full[1][k]) <= 1e-5 * max(1.0, abs(full[1][k])), (k, stats[k], full[1][k])
lo, hi = shard_range(text.shape[0], rank, world)
_, _, imv_s, _, _ = dp(text, tl, speech, sl)
assert torch.equal(imv_s, full[2][lo:hi])
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_data_parallel_forward_matches_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    port = 29600 + os.getpid() % 300
    code = WORKER % dict(root=ROOT)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out[-3000:]
        assert "ok" in out
