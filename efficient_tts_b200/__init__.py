"""efficient_tts_b200 -- the EFTS-CNN forward path of liusongxiang/efficient_tts, rebuilt as
hand-written sm_100a CUDA kernels behind a C ABI (``include/efts_b200.h``).

``efficient_tts_b200.models.EfficientTTSCNN`` mirrors ``nntts.models.EfficientTTSCNN``;
``efficient_tts_b200.layers`` mirrors the ``nntts.layers`` modules on the path;
``efficient_tts_b200.vocoder.Generator`` mirrors ``nntts.vocoders.hifigan_model.Generator`` (the HiFi-GAN
generator the reference runs on ``inference()``'s output).
"""
from . import models  # noqa: F401
from . import vocoder  # noqa: F401
from .layers import DurationPredictor, LengthRegulator, ResConvBlock  # noqa: F401
from .losses import FastSpeechLoss  # noqa: F401
from .models import EfficientTTSCNN  # noqa: F401

__all__ = ["EfficientTTSCNN", "ResConvBlock", "DurationPredictor", "LengthRegulator", "FastSpeechLoss", "models",
           "vocoder"]
