"""Training slice: ResConvBlock forward + backward on the GPU against torch autograd on the CPU
oracle (fp32).

Tolerances.  Forward as everywhere (1e-4 abs).  LeakyReLU' is discontinuous at 0: wherever a pre-activation lies
within the two implementations' fp32 noise of zero (a handful of the ~10^6 elements of a test) the two backward
passes legitimately take different slopes for that element, which moves single gradient entries by O(|g| |w|).  So
  * against torch autograd the gradients are compared in relative L2 norm (<= 1e-3; measured ~1e-6 without a flip), and
  * the exact check -- max error <= 5e-5 of the largest entry -- is made against a float64 evaluation of the same
    backward formulas that uses the slopes the device forward actually took (test_backward_matches_float64_...)."""
import numpy as np
import pytest
import torch

from oracle import efts_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def reference_grads(state, n_layers, x_bct, grad_bct):
    """torch autograd through the oracle's residual stack with the weight-norm parametrisation as leaves."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    x = x_bct.clone().requires_grad_(True)
    y = orc.res_conv_stack(x, {"blk." + k: v for k, v in leaves.items()}, "blk", n_layers)
    y.backward(grad_bct)
    return y.detach(), x.grad, {k: v.grad for k, v in leaves.items()}


@pytest.mark.parametrize("n_layers,B,T,seed", [(2, 2, 150, 0), (6, 3, 300, 1), (3, 1, 37, 2)])
def test_resconvblock_trains_like_the_reference(dev, n_layers, B, T, seed):
    from efficient_tts_b200.layers import ResConvBlock
    torch.manual_seed(seed)
    blk = ResConvBlock(n_layers, dropout_rate=0.0)
    state = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.randn(B, 512, T, generator=g)
    grad = torch.randn(B, 512, T, generator=g)
    y_ref, gx_ref, gp_ref = reference_grads(state, n_layers, x, grad)

    blk = blk.to(dev).train()
    xd = x.to(dev).requires_grad_(True)
    y = blk(xd)
    assert y.shape == y_ref.shape and y.requires_grad
    y.backward(grad.to(dev))
    assert (y.detach().cpu() - y_ref).abs().max().item() <= 1e-4

    def rel(a, b):
        return (a.cpu() - b).norm().item() / max(b.norm().item(), 1e-12)
    r = rel(xd.grad, gx_ref)
    print("dL/dx: relative L2 error %.2e" % r)
    assert r <= 1e-3
    params = dict(blk.named_parameters())
    assert set(params) == set(gp_ref)
    for k in sorted(params):
        assert params[k].grad is not None, k
        r = rel(params[k].grad, gp_ref[k])
        print("%s: relative L2 error %.2e" % (k, r))
        assert r <= 1e-3, k
    # the forward of the training path against the forward of the inference path: the only difference is where
    # the weight-norm fold g * v / ||v|| is evaluated (device vs host, one ulp on some weights)
    blk.eval()
    with torch.no_grad():
        y_inf = blk(x.to(dev))
    assert (y_inf - y.detach()).abs().max().item() <= 1e-5
    # without the reparametrisation both paths see the same weights: bit-identical results
    plain = ResConvBlock(2, dropout_rate=0.0, use_weight_norm=False).to(dev)
    y_train = plain.train()(x.to(dev).requires_grad_(True)).detach()
    with torch.no_grad():
        y_eval = plain.eval()(x.to(dev))
    assert torch.equal(y_train, y_eval)


@pytest.mark.parametrize("n_layers,B,T,k", [(3, 2, 200, 5), (2, 3, 131, 3), (1, 1, 40, 5)])
def test_backward_matches_float64_formulas_with_the_device_slopes(dev, n_layers, B, T, k):
    """dL/dx = g + conv^T(G'), dL/dW = sum_pos G' x_shifted, dL/db = sum_pos G' with G' = g * slope, evaluated in
    float64 on the CPU from the activations and slopes the device forward produced: max error <= 5e-5 of the largest
    entry of each gradient."""
    import torch.nn.functional as F
    from efficient_tts_b200.engine import train_context
    tc = train_context(dev)
    g = torch.Generator().manual_seed(40 + T)
    C = 512
    x = torch.randn(B, T, C, generator=g)
    w = torch.randn(n_layers, C, C, k, generator=g) / np.sqrt(C * k)
    b = torch.randn(n_layers, C, generator=g) * 0.1
    grad = torch.randn(B, T, C, generator=g)
    acts, us = tc.resconv_fwd(x.to(dev), w.to(dev), b.to(dev))
    gx, gw, gb = tc.resconv_bwd(grad.to(dev), acts, us, w.to(dev))
    acts64, us64 = acts.double().cpu(), us.double().cpu()
    # forward consistency of what was saved: acts[l + 1] = acts[l] + us[l], us[l] = lrelu(conv(acts[l]) + b)
    for l in range(n_layers):
        pre = F.conv1d(acts64[l].transpose(1, 2), w[l].double(), b[l].double(), padding=(k - 1) // 2).transpose(1, 2)
        assert (F.leaky_relu(pre, 0.1) - us64[l]).abs().max().item() <= 2e-5
        assert torch.equal(acts[l + 1], acts[l] + us[l])
    gcur = grad.double()
    for l in reversed(range(n_layers)):
        gp = gcur * torch.where(us64[l] > 0, 1.0, 0.1)                       # the slopes the device took
        gp_ct, x_ct = gp.transpose(1, 2), acts64[l].transpose(1, 2)
        dw = torch.nn.grad.conv1d_weight(x_ct, (C, C, k), gp_ct, padding=(k - 1) // 2)
        db = gp.sum((0, 1))
        gcur = gcur + F.conv_transpose1d(gp_ct, w[l].double(), padding=(k - 1) // 2).transpose(1, 2)
        for name, got, want in (("dW", gw[l], dw), ("db", gb[l], db)):
            err = (got.double().cpu() - want).abs().max().item() / want.abs().max().item()
            print("layer %d %s: max err %.2e of the largest entry" % (l, name, err))
            assert err <= 5e-5, (l, name)
    err = (gx.double().cpu() - gcur).abs().max().item() / gcur.abs().max().item()
    print("dL/dx: max err %.2e of the largest entry" % err)
    assert err <= 5e-5


def test_one_sgd_step_follows_the_reference(dev):
    """loss.backward() + optimiser step (trainers/efficient_tts_trainer.py:152-160 on the block alone): after one SGD
    step the parameters agree with the reference's step, and the loss goes down on the next forward."""
    from efficient_tts_b200.layers import ResConvBlock
    torch.manual_seed(5)
    blk = ResConvBlock(3, dropout_rate=0.0)
    state = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    g = torch.Generator().manual_seed(6)
    x, target = torch.randn(2, 512, 96, generator=g), torch.randn(2, 512, 96, generator=g)
    # reference step on the CPU
    leaves = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    loss_ref = ((orc.res_conv_stack(x, {"blk." + k: v for k, v in leaves.items()}, "blk", 3) - target) ** 2).mean()
    loss_ref.backward()
    stepped = {k: (v - 0.05 * v.grad).detach() for k, v in leaves.items()}
    # the same step through the library
    blk = blk.to(dev).train()
    opt = torch.optim.SGD(blk.parameters(), lr=0.05)
    loss = ((blk(x.to(dev)) - target.to(dev)) ** 2).mean()
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * max(1.0, abs(loss_ref.item()))
    for k, v in blk.state_dict().items():
        assert (v.cpu() - stepped[k]).abs().max().item() <= 1e-6 + 1e-4 * (stepped[k] - state[k]).abs().max().item(), k
    loss2 = ((blk(x.to(dev)) - target.to(dev)) ** 2).mean()
    assert loss2.item() < loss.item()


def test_weight_gradient_at_training_size_is_linear_and_matches_a_sample(dev):
    """Mid-size check (B = 16, T = 800: 12 800 positions per output element, several split items per tile): the weight
    gradient is linear in grad_out (exactly additive inputs -> additive outputs up to fp32 rounding) and a random
    sample of its entries matches a float64 evaluation of the defining sum."""
    from efficient_tts_b200.engine import train_context
    tc = train_context(dev)
    g = torch.Generator().manual_seed(9)
    B, T, C, k = 16, 800, 512, 5
    x = torch.randn(B, T, C, generator=g).to(dev)
    w = (torch.randn(1, C, C, k, generator=g) / np.sqrt(C * k)).to(dev)
    b = (torch.randn(1, C, generator=g) * 0.1).to(dev)
    acts, us = tc.resconv_fwd(x, w, b)
    g1 = torch.randn(B, T, C, generator=g).to(dev)
    g2 = torch.randn(B, T, C, generator=g).to(dev)
    _, gw1, gb1 = tc.resconv_bwd(g1, acts, us, w)
    _, gw2, gb2 = tc.resconv_bwd(g2, acts, us, w)
    _, gw12, gb12 = tc.resconv_bwd(g1 + g2, acts, us, w)
    scale = gw12.abs().max().item()
    assert (gw1 + gw2 - gw12).abs().max().item() <= 2e-5 * scale
    assert (gb1 + gb2 - gb12).abs().max().item() <= 2e-5 * gb12.abs().max().item()
    # float64 evaluation of dW[o, c, j] = sum_{b,t} G'[b,t,o] x[b, t + j - 2, c] on sampled (o, c, j)
    gp = (g1.double() * torch.where(us[0] > 0, 1.0, 0.1).double()).cpu()
    xp = torch.nn.functional.pad(x.double().cpu(), (0, 0, 2, 2))
    rng = np.random.default_rng(0)
    for _ in range(12):
        o, c, j = int(rng.integers(C)), int(rng.integers(C)), int(rng.integers(k))
        want = (gp[:, :, o] * xp[:, j:j + T, c]).sum().item()
        got = gw1[0, o, c, j].item()
        assert abs(got - want) <= 2e-4 * scale, (o, c, j, got, want)


# ---------------------------------------------------------------- data-parallel training (bin/train.py:210-216)
DDP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from efficient_tts_b200.layers import ResConvBlock
from torch.nn.parallel import DistributedDataParallel as DDP
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(7)                                   # same initial weights on every rank, like the reference's DDP
blk = ResConvBlock(3, dropout_rate=0.0).to(dev).train()
g = torch.Generator().manual_seed(11)
B, T = 4 * world, 173
x = torch.randn(B, 512, T, generator=g)
target = torch.randn(B, 512, T, generator=g)
lo, hi = rank * B // world, (rank + 1) * B // world
ddp = DDP(blk, device_ids=[rank])
opt = torch.optim.SGD(ddp.parameters(), lr=0.05)
y = ddp(x[lo:hi].to(dev))
loss = torch.nn.functional.mse_loss(y, target[lo:hi].to(dev))          # mean over the shard; DDP averages the gradients
loss.backward()
grads = {k: p.grad.detach().clone() for k, p in blk.named_parameters()}
# every rank holds the same (all-reduced) gradients
for k, v in grads.items():
    ref = v.clone(); dist.broadcast(ref, 0)
    assert torch.equal(v, ref), "rank %%d: gradient of %%s differs from rank 0 after the all-reduce" %% (rank, k)
# ... and they are the single-process gradients of the mean loss over the whole batch
torch.manual_seed(7)
solo = ResConvBlock(3, dropout_rate=0.0).to(dev).train()
l2 = torch.nn.functional.mse_loss(solo(x.to(dev)), target.to(dev))
l2.backward()
for k, p in solo.named_parameters():
    err = (grads[k] - p.grad).norm().item() / max(p.grad.norm().item(), 1e-20)
    assert err <= 1e-4, (k, err)
opt.step()
w0 = {k: p.detach().clone() for k, p in blk.named_parameters()}
for k, v in w0.items():
    ref = v.clone(); dist.broadcast(ref, 0)
    assert torch.equal(v, ref), "rank %%d: parameter %%s diverged after the step" %% (rank, k)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_resconvblock_trains_under_distributed_data_parallel():
    """The reference trains under torch DDP (bin/train.py:210-216): the gradient all-reduce is torch's bucketed NCCL
    all-reduce on the parameters' autograd hooks.  The drop-in block produces its gradients through a custom
    autograd.Function, so the same wrapper applies unchanged: two ranks, half a batch each, equal all-reduced
    gradients that match the single-process gradients of the whole batch, equal parameters after one SGD step."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world = 2
    port = 29900 + os.getpid() % 90
    code = DDP_WORKER % dict(root=root)
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


# ---------------------------------------------------------------- duration predictor (layers/duration_predictor.py)
def _dp_inputs(B, T, seed, dev):
    g = torch.Generator().manual_seed(seed)
    xs = torch.randn(B, T, 512, generator=g)
    lengths = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
    lengths[0] = T
    masks = torch.arange(T)[None, :] >= lengths[:, None]            # True = padded (x_masks of the reference)
    grad = torch.randn(B, T, generator=g)
    return xs, masks, grad


@pytest.mark.parametrize("B,T,seed", [(3, 77, 0), (2, 200, 1)])
def test_duration_predictor_trains_like_the_reference(dev, B, T, seed):
    """DurationPredictor.forward under autograd (log-domain output, masked positions 0) against torch autograd through the
    CPU oracle: output <= 1e-4 abs; gradients of every parameter and of the input in relative L2 (ReLU' is discontinuous
    at 0, see the module docstring)."""
    from efficient_tts_b200.layers import DurationPredictor
    torch.manual_seed(seed)
    dp = DurationPredictor(512, n_layers=2, n_chans=512, kernel_size=3, dropout_rate=0.0)
    with torch.no_grad():                                  # non-trivial affine parameters
        for seq in dp.conv:
            seq[2].weight.uniform_(0.5, 1.5)
            seq[2].bias.uniform_(-0.3, 0.3)
    xs, masks, grad = _dp_inputs(B, T, 40 + seed, dev)
    leaves = {"duration_predictor." + k: v.detach().clone().requires_grad_(True) for k, v in dp.state_dict().items()}
    x_ref = xs.clone().requires_grad_(True)
    out_ref = orc.duration_predictor_forward(x_ref, masks, leaves)
    out_ref.backward(grad)

    dp = dp.to(dev).train()
    xd = xs.to(dev).requires_grad_(True)
    out = dp(xd, masks.to(dev))
    assert out.shape == out_ref.shape and out.requires_grad
    out.backward(grad.to(dev))
    assert (out.detach().cpu() - out_ref.detach()).abs().max().item() <= 1e-4
    assert torch.all(out.detach().cpu()[masks] == 0)

    def rel(a, b):
        return (a.cpu() - b).norm().item() / max(b.norm().item(), 1e-12)
    r = rel(xd.grad, x_ref.grad)
    print("dL/dxs: relative L2 error %.2e" % r)
    assert r <= 1e-3
    for k, p in dp.named_parameters():
        assert p.grad is not None, k
        r = rel(p.grad, leaves["duration_predictor." + k].grad)
        print("%s: relative L2 error %.2e" % (k, r))
        assert r <= 1e-3, k
    # eval-mode forward of the same module (the inference kernels) agrees with the training forward
    dp.eval()
    with torch.no_grad():
        out_eval = dp(xs.to(dev), masks.to(dev))
    assert (out_eval - out.detach()).abs().max().item() <= 2e-5


def test_duration_predictor_backward_matches_float64_with_dropout_masks(dev):
    """The raw library pair with train-mode dropout masks, against a float64 evaluation that applies the same masks and the
    ReLU slopes the device forward took: max error <= 5e-5 of the largest entry of every gradient."""
    from efficient_tts_b200.engine import train_context
    L, B, T, C, k = 2, 2, 131, 512, 3
    g = torch.Generator().manual_seed(5)
    xs, masks, grad = _dp_inputs(B, T, 77, dev)
    conv_w = torch.randn(L, C, C, k, generator=g) / np.sqrt(C * k)
    conv_b = torch.randn(L, C, generator=g) * 0.1
    ln_g = torch.rand(L, C, generator=g) + 0.5
    ln_b = torch.randn(L, C, generator=g) * 0.2
    head_w = torch.randn(C, generator=g) / np.sqrt(C)
    head_b = torch.randn(1, generator=g)
    keep = (torch.rand(L, B, T, C, generator=g) >= 0.1).float() / 0.9
    tc = train_context(dev)
    d = lambda t: t.to(dev)
    out, acts, us = tc.duration_fwd(d(xs), d(conv_w), d(conv_b), d(ln_g), d(ln_b), d(head_w), d(head_b), d(masks), d(keep))
    got = tc.duration_bwd(d(grad), acts, us, d(conv_w), d(ln_g), d(head_w), d(masks), d(keep))

    leaves = [t.double().clone().requires_grad_(True) for t in (xs, conv_w, conv_b, ln_g, ln_b, head_w, head_b)]
    x64, w64, b64, g64, be64, hw64, hb64 = leaves
    h = x64
    for l in range(L):
        pre = torch.nn.functional.conv1d(h.transpose(1, 2), w64[l], b64[l], padding=(k - 1) // 2).transpose(1, 2)
        u = pre * (us[l].cpu() > 0).double()                        # the slopes the device took
        assert (torch.relu(pre).float() - us[l].cpu()).abs().max().item() <= 2e-5
        h = torch.nn.functional.layer_norm(u, (C,), g64[l], be64[l], eps=1e-12) * keep[l].double()
    o64 = (h @ hw64 + hb64).masked_fill(masks, 0.0)
    assert (out.cpu().double() - o64.detach()).abs().max().item() <= 1e-4
    o64.backward(grad.double())
    names = ["x", "conv_w", "conv_b", "ln_g", "ln_b", "head_w", "head_b"]
    for name, a, leaf in zip(names, got, leaves):
        want = leaf.grad
        err = (a.cpu().double().reshape(want.shape) - want).abs().max().item() / max(want.abs().max().item(), 1e-30)
        print("%s: max error / max |grad| = %.2e" % (name, err))
        assert err <= 5e-5, name


# ---------------------------------------------------------------- criterion (losses/fastspeech_loss.py)
def test_fastspeech_loss_module_matches_the_reference_fixture_and_the_oracle(dev):
    """efficient_tts_b200.losses.FastSpeechLoss: values and gradients against the fixture of the unmodified reference
    module (tests/golden/loss_cases.npz) and, at a training-sized shape, against autograd through the oracle."""
    import os
    from efficient_tts_b200.losses import FastSpeechLoss
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_cases.npz"))
    for name in ("mask_mse", "nomask_mse", "mask_l1"):
        seed, B, T1, T2, odim, use_masking, use_mse = (int(v) for v in z[name + ".cfg"])
        mel, dur, ys, ds, il, ol = (t.to(dev) for t in orc.make_loss_inputs(seed, B, T1, T2, odim))
        mel.requires_grad_(True)
        dur.requires_grad_(True)
        crit = FastSpeechLoss(use_masking=bool(use_masking), use_mse=bool(use_mse))
        ml, dl = crit(None, mel, dur, ys, ds, il, ol)
        (ml * 0.7 + dl * 1.3).backward()
        want = z[name + ".losses"]
        assert abs(ml.item() - want[0]) <= 2e-6 * max(1.0, abs(want[0])), (name, ml.item(), want[0])
        assert abs(dl.item() - want[1]) <= 2e-6 * max(1.0, abs(want[1])), (name, dl.item(), want[1])
        gm, gd = z[name + ".grad_mel"], z[name + ".grad_dur"]
        assert np.abs(mel.grad.cpu().numpy() - gm).max() <= 1e-6 * np.abs(gm).max(), name
        assert np.abs(dur.grad.cpu().numpy() - gd).max() <= 1e-6 * np.abs(gd).max(), name
    # training-sized: B = 32, 200 tokens -> 1200 frames
    mel, dur, ys, ds, il, ol = orc.make_loss_inputs(9, 32, 200, 1200, 80)
    mr, dr = mel.clone().requires_grad_(True), dur.clone().requires_grad_(True)
    a, b = orc.fastspeech_loss(mr, dr, ys, ds, il, ol, True, True)
    (a + b).backward()
    md, dd = mel.to(dev).requires_grad_(True), dur.to(dev).requires_grad_(True)
    ml, dl = FastSpeechLoss()(None, md, dd, ys.to(dev), ds.to(dev), il.to(dev), ol.to(dev))
    (ml + dl).backward()
    assert abs(ml.item() - a.item()) <= 2e-6 * abs(a.item()) and abs(dl.item() - b.item()) <= 2e-6 * abs(b.item())
    assert (md.grad.cpu() - mr.grad).abs().max().item() <= 1e-6 * mr.grad.abs().max().item()
    assert (dd.grad.cpu() - dr.grad).abs().max().item() <= 1e-6 * dr.grad.abs().max().item()
    # the reference's mask rule: padded dims must equal the longest length
    with pytest.raises(RuntimeError):
        FastSpeechLoss()(None, md, dd, ys.to(dev), ds.to(dev), il.to(dev), (ol - 1).clamp(min=1).to(dev))


def test_duration_predictor_training_step_with_the_criterion(dev):
    """One optimiser step of the duration branch as the reference runs it (models/efficient_tts.py:219-221:
    dur_pred = duration_predictor(text_value, ~text_mask); criterion(..., dur_pred, ..., log_delta_e, ...)), every
    gradient from the library: parameters after the step match the CPU reference within 1e-6 + 1e-4 of the update."""
    from efficient_tts_b200.layers import DurationPredictor
    from efficient_tts_b200.losses import FastSpeechLoss
    torch.manual_seed(3)
    dp = DurationPredictor(512, n_layers=2, n_chans=512, kernel_size=3, dropout_rate=0.0)
    state = {k: v.detach().clone() for k, v in dp.state_dict().items()}
    B, T1, T2, odim = 4, 60, 90, 80
    mel, _, ys, ds, il, ol = orc.make_loss_inputs(21, B, T1, T2, odim)
    xs = torch.randn(B, T1, 512, generator=torch.Generator().manual_seed(22))
    masks = torch.arange(T1)[None, :] >= il[:, None]
    lr = 0.05
    # CPU reference: oracle forward + torch autograd + SGD
    leaves = {"duration_predictor." + k: v.clone().requires_grad_(True) for k, v in state.items()}
    d_ref = orc.duration_predictor_forward(xs, masks, leaves)
    a, b = orc.fastspeech_loss(mel, d_ref, ys, ds, il, ol, True, True)
    (a + b).backward()
    stepped = {k: leaves["duration_predictor." + k].detach() - lr * leaves["duration_predictor." + k].grad for k in state}
    # device
    dp = dp.to(dev).train()
    opt = torch.optim.SGD(dp.parameters(), lr=lr)
    d_out = dp(xs.to(dev), masks.to(dev))
    ml, dl = FastSpeechLoss()(None, mel.to(dev), d_out, ys.to(dev), ds.to(dev), il.to(dev), ol.to(dev))
    loss = ml + dl
    loss.backward()
    opt.step()
    assert abs(loss.item() - (a + b).item()) <= 1e-5 * max(1.0, abs((a + b).item()))
    for k, v in dp.state_dict().items():
        upd = (stepped[k] - state[k]).abs().max().item()
        assert (v.cpu() - stepped[k]).abs().max().item() <= 1e-6 + 1e-4 * upd, k
