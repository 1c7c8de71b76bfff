"""Drop-in for ``nntts.models.EfficientTTSCNN`` (``/root/reference/nntts/models/efficient_tts.py``).

The reference resolves its model by name -- ``getattr(nntts.models, config["model_name"])(**params)``
(bin/train.py:173-185, bin/inference.py:63-75) -- then uses ``load_state_dict``,
``remove_weight_norm()``, ``.eval().to(device)``, ``model(text=, text_lengths=, speech=,
speech_lengths=)`` and ``model.inference(text)``.  This class keeps that surface (constructor
kwargs, parameter names, return tuples) and runs the forward path in ``libefts_b200.so``.
"""
import torch

from .layers import DurationPredictor, ResConvBlock, _EngineOwner


def raise_on_flags(flags, check_padded_dims=False):
    """Turn the device error word of a forward pass (include/efts_b200.h, ``scalars[7]``) into the exception the
    reference raises for the same input.  ``check_padded_dims`` applies the max(lengths) == padded-dim rule of the
    reference's mask builder, which a data-parallel shard with global padded dims is exempt from."""
    if flags & 4:
        raise IndexError("index out of range in self")          # torch.nn.Embedding, models/efficient_tts.py:144
    if flags & 8:
        from .engine import RANGE_MESSAGE
        raise FloatingPointError(RANGE_MESSAGE + " [flags 0x%x]" % flags)
    if flags & 16:
        raise RuntimeError("a length lies outside [0, padded dim] (the reference's mask broadcast fails on such "
                           "input, utils/nets_utils.py:148-168)")
    if check_padded_dims and flags & 3:
        which = "text" if flags & 1 else "speech"
        raise RuntimeError("The padded %s length must equal max(%s_lengths) (the reference builds its "
                           "masks with maxlen = max(lengths), utils/nets_utils.py:148)" % (which, which))


class EfficientTTSCNN(_EngineOwner):
    """EFTS-CNN, forward path on B200.  Constructor: models/efficient_tts.py:26-49."""

    def __init__(self, num_symbols, odim=80, symbol_embedding_dim=512, n_channels=512,
                 n_text_encoder_layer=5, n_mel_encoder_layer=3, n_decoder_layer=6, n_duration_layer=2,
                 k_size=5, nonlinear_activation="LeakyReLU",
                 nonlinear_activation_params={"negative_slope": 0.1}, use_weight_norm=True,
                 dropout_rate=0.1, use_masking=False, use_weighted_masking=False, duration_offset=1.0,
                 sigma=0.01, sigma_e=0.5, delta_e_method_1=True, share_text_encoder_key_value=False,
                 use_mel_query_fc=False):
        super().__init__()
        # configurations outside the production recipe are rejected, never approximated
        if share_text_encoder_key_value:
            raise NotImplementedError("share_text_encoder_key_value=True is outside the B200 path")
        if use_mel_query_fc:
            raise NotImplementedError("use_mel_query_fc=True is outside the B200 path")
        if not delta_e_method_1:
            raise NotImplementedError("delta_e_method_1=False is outside the B200 path")
        if use_weighted_masking:
            raise NotImplementedError("use_weighted_masking=True is outside the B200 path (the LJ recipe, "
                                      "egs/lj/conf/efficient_tts_cnn_phnseq_noDropout.v1.yaml:21, disables it)")
        if symbol_embedding_dim != n_channels:
            raise NotImplementedError("symbol_embedding_dim must equal n_channels")
        self.num_symbols, self.odim, self.n_channels, self.k_size = num_symbols, odim, n_channels, k_size
        self.use_masking = bool(use_masking)
        self.duration_offset = duration_offset
        self.sigma = sigma
        self.sigma_e = sigma_e
        self.delta_e_method_1 = delta_e_method_1
        self.share_text_encoder_key_value = share_text_encoder_key_value
        blk = dict(n_channels=n_channels, k_size=k_size, nonlinear_activation=nonlinear_activation,
                   nonlinear_activation_params=nonlinear_activation_params, dropout_rate=dropout_rate,
                   use_weight_norm=use_weight_norm)
        # same construction order as the reference, so a seeded default init gives the same weights
        self.text_embedding_table = torch.nn.Embedding(num_symbols, symbol_embedding_dim)
        self.text_encoder = ResConvBlock(num_layers=n_text_encoder_layer, **blk)
        self.text_encoder_key = torch.nn.Linear(n_channels, n_channels)
        self.text_encoder_value = torch.nn.Linear(n_channels, n_channels)
        self.mel_prenet = torch.nn.Sequential(
            torch.nn.Linear(odim, n_channels),
            getattr(torch.nn, nonlinear_activation)(**nonlinear_activation_params),
            torch.nn.Dropout(dropout_rate))
        self.mel_encoder = ResConvBlock(num_layers=n_mel_encoder_layer, **blk)
        self.mel_query_fc = None
        self.decoder = ResConvBlock(num_layers=n_decoder_layer, **blk)
        self.mel_output_layer = torch.nn.Linear(n_channels, odim)
        self.duration_predictor = DurationPredictor(idim=n_channels, n_layers=n_duration_layer,
                                                    n_chans=n_channels, offset=duration_offset)

    # ------------------------------------------------------------------ engine plumbing
    def _engine_kwargs(self):
        return dict(num_symbols=self.num_symbols, odim=self.odim, n_channels=self.n_channels,
                    k_size=self.k_size, n_text_encoder_layer=self.text_encoder.num_layers,
                    n_mel_encoder_layer=self.mel_encoder.num_layers, n_decoder_layer=self.decoder.num_layers,
                    n_duration_layer=self.duration_predictor.n_layers,
                    duration_kernel_size=self.duration_predictor.kernel_size, sigma=self.sigma,
                    sigma_e=self.sigma_e, duration_offset=self.duration_offset,
                    use_masking=self.use_masking)

    def _engine_state(self):
        return self.state_dict()

    def _get_engine(self):
        eng = super()._get_engine()
        # sub-modules share this context's fingerprinted lifetime; nothing else to do
        return eng

    # ------------------------------------------------------------------ reference surface
    def forward(self, text, text_lengths, speech, speech_lengths):
        """models/efficient_tts.py:120-228 ->
        ``(loss, stats, imv, reconst_alpha, mel_pred, speech)``."""
        self._require_eval()
        eng = self._get_engine()
        imv, reconst_alpha, mel_pred, scal = eng.forward(text, text_lengths, speech, speech_lengths)
        host = scal.cpu()                      # the reference's three .item() syncs (:225-227) in one
        raise_on_flags(int(host[7]), check_padded_dims=True)
        stats = dict(loss=float(host[0]), mel_loss=float(host[1]), duration_loss=float(host[2]))
        return scal[0], stats, imv, reconst_alpha, mel_pred, speech

    def forward_shard(self, text, text_lengths, speech, speech_lengths):
        """One data-parallel shard of a larger batch: the tensors keep the GLOBAL padded dims, so the
        max(lengths) == padded-dim check of ``forward`` does not apply; nothing is read back.
        Returns ``(imv, reconst_alpha, mel_pred, scalars[8])`` on the device (include/efts_b200.h)."""
        self._require_eval()
        return self._get_engine().forward(text, text_lengths, speech, speech_lengths)

    def inference(self, text, text_lengths=None):
        """models/efficient_tts.py:230-285 -> ``(mel_pred[1,T2,odim], reconst_alpha[1,T1,T2])``."""
        self._require_eval()
        return self._get_engine().inference(text)

    def inference_batch(self, text, text_lengths):
        """Batched variable-length synthesis (the reference stops at B = 1 because of the
        ``.item()`` at models/efficient_tts.py:361).  ``text`` int64 [B, T1] (padded anyhow), ``text_lengths``
        [B].  Returns ``(mel_pred[B, T2max, odim], mel_lengths[B], reconst_alpha[B, T1, T2max])`` where row b
        equals ``inference(text[b:b+1, :text_lengths[b]])`` and is zero beyond its own length."""
        self._require_eval()
        return self._get_engine().inference_batch(text, text_lengths)

    # ------------------------------------------------------------------ the reference's helper methods
    # (models/efficient_tts.py:287-398).  forward() runs fused kernels; these stand-alone forms keep the
    # public method surface.  Masks are the prefix masks make_non_pad_mask builds; they travel as lengths.
    def generate_index_vector(self, text_mask):
        from . import engine as E
        return E.index_vector(E.mask_lengths(text_mask), text_mask.shape[1])

    def imv_generator(self, alpha, p, mel_mask, text_length):
        from . import engine as E
        return E.imv_generator(alpha, p, text_length.to(alpha.device, torch.int32).contiguous(), E.mask_lengths(mel_mask))

    def get_aligned_positions(self, imv, p, mel_mask, text_mask, sigma=0.5):
        from . import engine as E
        return E.aligned_positions(imv, p, E.mask_lengths(text_mask), E.mask_lengths(mel_mask), sigma)

    def reconstruct_align_from_aligned_position(self, e, delta=0.1, mel_mask=None, text_mask=None, trim_e=False):
        from . import engine as E
        if trim_e:
            raise NotImplementedError("trim_e is only reachable with delta_e_method_1=False (outside the B200 path)")
        if mel_mask is None:
            max_length = int(torch.round(e[:, -1]).squeeze().item())       # the reference's .item(), :361
            sl = None
        else:
            max_length = mel_mask.size(-1)
            sl = E.mask_lengths(mel_mask)
        tl = E.mask_lengths(text_mask) if text_mask is not None else None
        return E.reconstruct_alignment(e, delta, tl, sl, max_length)

    def scaled_dot_product_attention(self, query, key, key_mask):
        from . import engine as E
        return self._get_engine().attention_alpha(query, key, E.mask_lengths(key_mask))

    def remove_weight_norm(self):
        for m in (self.text_encoder, self.mel_encoder, self.decoder):
            m.remove_weight_norm()

    def apply_weight_norm(self):
        for m in (self.text_encoder, self.mel_encoder, self.decoder):
            m.apply_weight_norm()
