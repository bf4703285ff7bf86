"""Host side of the GPU resampler: the polyphase windowed-sinc filter table of torchaudio.transforms.Resample
(default `sinc_interp_hann`, lowpass_filter_width 6, rolloff 0.99), restated from
torchaudio/functional/functional.py `_get_sinc_resample_kernel` (the reference pins torchaudio==2.2.2,
asr/requirements.txt:1; call site asr/wenet/cli/reverb.py:125-128).  The convolution itself runs on the GPU
(csrc/resample.cu, `rvb_resample`)."""
import functools
import math

import numpy as np

LOWPASS_FILTER_WIDTH = 6
ROLLOFF = 0.99


@functools.lru_cache(maxsize=16)
def sinc_resample_kernel(orig_freq: int, new_freq: int):
    """-> (kernel (new, 2*width+orig) float32, orig, new, width) with orig/new reduced by their gcd."""
    if not (int(orig_freq) == orig_freq and int(new_freq) == new_freq):
        raise Exception("Frequencies must be of integer type to ensure quality resampling computation.")
    gcd = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // gcd, int(new_freq) // gcd
    base_freq = min(orig, new) * ROLLOFF
    width = math.ceil(LOWPASS_FILTER_WIDTH * orig / base_freq)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    # torchaudio divides an int64 arange by a Python int: a float32 division, promoted to float64 by the addition
    phase = (np.arange(0, -new, -1, dtype=np.int64).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None]
    t = (phase + idx) * base_freq
    t = np.clip(t, -LOWPASS_FILTER_WIDTH, LOWPASS_FILTER_WIDTH)
    window = np.cos(t * math.pi / LOWPASS_FILTER_WIDTH / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    with np.errstate(divide="ignore", invalid="ignore"):
        kern = np.where(t == 0, 1.0, np.sin(t) / t)
    kern = (kern * window * scale).astype(np.float32)
    return np.ascontiguousarray(kern), orig, new, width


def resampled_length(n_in: int, orig: int, new: int) -> int:
    return int(math.ceil(new * n_in / orig))
