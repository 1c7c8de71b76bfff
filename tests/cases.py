"""Seeded synthetic inputs for the parity tests and golden fixtures (see
``efficient_tts_b200/workloads.py``, which bench.py shares)."""
from efficient_tts_b200.workloads import (NUM_SYMBOLS, ODIM, config_lengths,  # noqa: F401
                                          make_forward_inputs, make_inference_inputs)
