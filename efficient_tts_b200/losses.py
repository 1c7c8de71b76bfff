"""The criterion of the path as a module: ``nntts.losses.fastspeech_loss.FastSpeechLoss`` (losses/fastspeech_loss.py:8-67).

Inside ``EfficientTTSCNN.forward`` the two terms are computed by the forward kernels (loss_partial / loss_finalize);
this mirror is the stand-alone, DIFFERENTIABLE form: one library call returns both terms and their
gradients with respect to the predictions, so ``mel_loss + duration_loss`` can be back-propagated into the modules that
train on this path (``DurationPredictor``, ``ResConvBlock``).  CUDA only; nothing computes on the CPU.
"""
import torch

from . import engine as _engine


class FastSpeechLoss(torch.nn.Module):
    """Constructor and ``forward`` signature of the reference (losses/fastspeech_loss.py:11-52)."""

    def __init__(self, use_masking: bool = True, use_weighted_masking: bool = False, use_mse: bool = True):
        super().__init__()
        assert (use_masking != use_weighted_masking) or not use_masking          # :25
        if use_weighted_masking:
            raise NotImplementedError("use_weighted_masking is outside the EFTS-CNN recipes "
                                      "(models/efficient_tts.py:115-117 passes the model's flags; egs/lj/conf/*.yaml "
                                      "leave it False)")
        self.use_masking = use_masking
        self.use_weighted_masking = False
        self.use_mse = use_mse

    def forward(self, after_outs, before_outs, d_outs, ys, ds, ilens, olens):
        if after_outs is not None:
            raise NotImplementedError("after_outs: the EFTS model has no postnet and passes None "
                                      "(models/efficient_tts.py:220)")
        if before_outs.device.type != "cuda":
            raise RuntimeError("efficient_tts_b200 modules compute on a CUDA sm_100a device only")
        if self.use_masking:
            # make_non_pad_mask builds the masks with maxlen = max(lengths) (utils/nets_utils.py:148): the reference's
            # masked_select only broadcasts when that equals the padded dim
            if int(olens.max()) != before_outs.shape[1] or int(ilens.max()) != d_outs.shape[1]:
                raise RuntimeError("The padded lengths must equal max(olens) / max(ilens) (the reference builds its masks "
                                   "with maxlen = max(lengths), utils/nets_utils.py:148)")
        return _engine.FastSpeechLossFunction.apply(before_outs, d_outs, ys, ds, ilens, olens, self.use_masking,
                                                    self.use_mse)
