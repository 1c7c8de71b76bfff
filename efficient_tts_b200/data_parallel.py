"""Data-parallel forward over utterance shards.

The reference's only parallelism is DDP with a ``DistributedSampler`` (bin/train.py:136-150,
210-216): one process per GPU, replicated weights, independent utterances.  The forward-only
analogue here: contiguous utterance ranges per rank, every rank keeping the caller's GLOBAL padded
(T1, T2) because padding is not inert, no collective on the data path.  NCCL is
used only for the two things that cross shards: the four loss partial sums (one all-reduce of a
4-float vector) and, on request, the concatenation of the per-rank outputs (all-gather).
"""
import torch
import torch.distributed as dist


ERROR_BITS = (2, 3, 4)     # scalars[7] bits that are errors on every path (include/efts_b200.h)


def shard_range(n_items, rank, world_size):
    """Contiguous [lo, hi) of ``n_items`` owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_loss_partials(partials):
    """``partials`` = (sum_sq, n_mel, sum_abs, n_tok) summed over ranks -> (loss, mel_loss,
    duration_loss): the same masked means FastSpeechLoss takes over the whole batch
    (losses/fastspeech_loss.py:54-67), models/efficient_tts.py:223."""
    sum_sq, n_mel, sum_abs, n_tok = (float(v) for v in partials)
    mel = sum_sq / n_mel
    dur = sum_abs / n_tok
    return mel + dur, mel, dur


def all_gather_rows(local, sizes, group=None):
    """Concatenate per-rank row blocks of different heights (rank r contributes ``sizes[r]`` rows)."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    hmax = max(sizes)
    pad = local.new_zeros((hmax,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


class DataParallelForward:
    """Runs ``forward`` of a replicated model on this rank's utterance shard.

    ``shard_forward(text, text_lengths, speech, speech_lengths) -> (imv, reconst_alpha, mel_pred,
    scalars[8])`` is ``EfficientTTSCNN.forward_shard`` in production; tests inject a CPU stand-in
    so the split / reduce / gather logic runs under gloo.
    """

    def __init__(self, shard_forward, group=None):
        self.shard_forward = shard_forward
        self.group = group

    def launch_local(self, text, text_lengths, speech, speech_lengths):
        """This rank's shard (tensors with the GLOBAL padded dims) -> ``(imv, reconst_alpha, mel_pred, part)``, all on
        the device and nothing read back: ``part`` is the all-reduced vector (sum_sq, n_mel, sum_abs, n_tok, bit
        counters).  One collective for everything that crosses shards: the four loss partial sums and, as 0/1
        counters, the error bits a single-process forward() raises on (bit 2 token id, bit 3 fp16 operand range,
        bit 4 length outside the padded dims) -- a rank that hit one must stop every rank, not just itself."""
        imv, ra, mel, scal = self.shard_forward(text, text_lengths, speech, speech_lengths)
        flags = scal[7].to(torch.int32)
        bits = torch.stack([(flags >> k) & 1 for k in ERROR_BITS]).to(scal.dtype)
        part = torch.cat([scal[3:7], bits])
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        return imv, ra, mel, part

    @staticmethod
    def finish(part):
        """Host side of ``launch_local``: reads the reduced vector (the reference's ``.item()`` on the loss), raises
        what ``forward()`` raises, returns ``(loss, stats)`` of the whole batch."""
        host = part.tolist()
        raised = sum((1 << k) for k, n in zip(ERROR_BITS, host[4:]) if n > 0)
        if raised:
            from .models import raise_on_flags
            raise_on_flags(raised)
        loss, mel_loss, dur_loss = combine_loss_partials(host[:4])
        return loss, dict(loss=loss, mel_loss=mel_loss, duration_loss=dur_loss)

    def __call__(self, text, text_lengths, speech, speech_lengths, gather_outputs=False):
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        B = text.shape[0]
        lo, hi = shard_range(B, rank, world)
        imv, ra, mel, part = self.launch_local(text[lo:hi], text_lengths[lo:hi], speech[lo:hi], speech_lengths[lo:hi])
        loss, stats = self.finish(part)
        if gather_outputs:
            sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
            imv = all_gather_rows(imv, sizes, self.group)
            ra = all_gather_rows(ra, sizes, self.group)
            mel = all_gather_rows(mel, sizes, self.group)
        return loss, stats, imv, ra, mel
