"""ORACLE (test infrastructure, never shipped): restatement of torchaudio.transforms.Resample as the reference calls it
(asr/wenet/cli/reverb.py:125-128: `Resample(orig_freq=sample_rate, new_freq=16000)(waveform)`, defaults
sinc_interp_hann / lowpass_filter_width 6 / rolloff 0.99).  The algorithm lives in the third-party dependency
torchaudio (pinned 2.2.2 in asr/requirements.txt:1; restated from torchaudio/functional/functional.py
`_get_sinc_resample_kernel` + `_apply_sinc_resample_kernel`, identical in 2.2.2 and the container's 2.11).
Pinned against torchaudio itself: oracle/make_golden_resample.py -> tests/golden/resample.npz."""
import math

import torch


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    gcd = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // gcd, int(new_freq) // gcd
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=None)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels.to(torch.float32), orig, new, width


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """waveform (C, N) float32 -> (C, ceil(N * new / orig))."""
    if orig_freq == new_freq:
        return waveform
    kernel, orig, new, width = sinc_resample_kernel(orig_freq, new_freq)
    num_wavs, length = waveform.shape
    x = torch.nn.functional.pad(waveform, (width, width + orig))
    y = torch.nn.functional.conv1d(x[:, None], kernel, stride=orig)
    y = y.transpose(1, 2).reshape(num_wavs, -1)
    target = int(torch.ceil(torch.as_tensor(new * length / orig)).long())
    return y[..., :target]
