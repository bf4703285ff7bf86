"""Alias module: `from wenet.cli.reverb import load_model, ReverbASR` (reference: asr/wenet/cli/reverb.py)."""
from reverb_b200.reverb import (ReverbASR, download_model, get_available_models, get_output,  # noqa: F401
                                load_model)
