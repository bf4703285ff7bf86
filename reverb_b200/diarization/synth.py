"""Synthetic weights for the two diarization networks, under pyannote's state_dict names.

The `Revai/reverb-diarization-v1` checkpoints live on HuggingFace and cannot be fetched offline, so
parity tests and the diarization benchmark run on seeded random weights of the published shapes.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

SEG_SHAPE = dict(sample_rate=16000, sinc_filters=80, sinc_kernel=251, sinc_stride=10, conv_channels=60, conv_kernel=5,
                 lstm_hidden=128, lstm_layers=4, linear_dim=128, linear_layers=2, num_classes=7)


def _uniform(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def segmentation_state_dict(seed: int = 0, shape: Dict = SEG_SHAPE) -> Dict[str, np.ndarray]:
    """PyanNet weights: SincNet cut-offs mel-spaced like ParamSincFB's initialisation (+ jitter), PyTorch-default
    uniform initialisation elsewhere; the classifier is scaled up so the log-probabilities are not flat."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, np.ndarray] = {}
    half = shape["sinc_filters"] // 2
    sr = shape["sample_rate"]

    def to_mel(hz):
        return 2595 * np.log10(1 + hz / 700)

    def to_hz(mel):
        return 700 * (10 ** (mel / 2595) - 1)

    mel = np.linspace(to_mel(30.0), to_mel(sr / 2 - 100.0), half + 1)
    hz = to_hz(mel)
    sd["sincnet.wav_norm1d.weight"] = np.array([1.1], np.float32)
    sd["sincnet.wav_norm1d.bias"] = np.array([0.02], np.float32)
    sd["sincnet.conv1d.0.filterbank.low_hz_"] = (hz[:-1] * rng.uniform(0.95, 1.05, half)).astype(np.float32).reshape(-1, 1)
    sd["sincnet.conv1d.0.filterbank.band_hz_"] = (np.diff(hz) * rng.uniform(0.9, 1.1, half)).astype(np.float32).reshape(-1, 1)
    chans = [shape["sinc_filters"], shape["conv_channels"], shape["conv_channels"]]
    for i, c in enumerate(chans):
        sd[f"sincnet.norm1d.{i}.weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        sd[f"sincnet.norm1d.{i}.bias"] = rng.uniform(-0.2, 0.2, c).astype(np.float32)
    k = shape["conv_kernel"]
    for i in (1, 2):
        cin = chans[i - 1]
        bound = 1.0 / math.sqrt(cin * k)
        sd[f"sincnet.conv1d.{i}.weight"] = _uniform(rng, (shape["conv_channels"], cin, k), bound)
        sd[f"sincnet.conv1d.{i}.bias"] = _uniform(rng, (shape["conv_channels"],), bound)
    H = shape["lstm_hidden"]
    bound = 1.0 / math.sqrt(H)
    for layer in range(shape["lstm_layers"]):
        cin = shape["conv_channels"] if layer == 0 else 2 * H
        for sfx in ("", "_reverse"):
            # larger than PyTorch's default so that the random network reacts to its input (default-init LSTMs are flat)
            sd[f"lstm.weight_ih_l{layer}{sfx}"] = _uniform(rng, (4 * H, cin), 4.0 * bound)
            sd[f"lstm.weight_hh_l{layer}{sfx}"] = _uniform(rng, (4 * H, H), 2.0 * bound)
            sd[f"lstm.bias_ih_l{layer}{sfx}"] = _uniform(rng, (4 * H,), bound)
            sd[f"lstm.bias_hh_l{layer}{sfx}"] = _uniform(rng, (4 * H,), bound)
    cin = 2 * H
    for i in range(shape["linear_layers"]):
        bound = 3.0 / math.sqrt(cin)
        sd[f"linear.{i}.weight"] = _uniform(rng, (shape["linear_dim"], cin), bound)
        sd[f"linear.{i}.bias"] = _uniform(rng, (shape["linear_dim"],), bound)
        cin = shape["linear_dim"]
    bound = 6.0 / math.sqrt(cin)
    sd["classifier.weight"] = _uniform(rng, (shape["num_classes"], cin), bound)
    sd["classifier.bias"] = _uniform(rng, (shape["num_classes"],), 0.5)
    return sd


def synthetic_speech(seconds: float, seed: int = 0, sample_rate: int = 16000, turns: int = 0) -> np.ndarray:
    """Speech-like float32 waveform in [-1, 1]: amplitude-modulated noise + harmonics; `turns` > 0 alternates between
    that many synthetic 'speakers' (different pitch / spectral tilt) every few seconds with short silences."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sample_rate)
    t = np.arange(n) / sample_rate
    out = np.zeros(n, np.float32)
    nspk = max(turns, 1)
    seg = 0
    pos = 0
    while pos < n:
        dur = int(rng.uniform(1.5, 4.0) * sample_rate)
        end = min(n, pos + dur)
        spk = seg % nspk
        f0 = 100.0 + 45.0 * spk
        tt = t[pos:end]
        env = 0.5 * (1 + np.sin(2 * np.pi * (3.0 + spk) * tt))
        sig = sum(np.sin(2 * np.pi * f0 * h * tt) / h for h in (1, 2, 3, 4))
        noise = rng.normal(0, 0.3, end - pos)
        out[pos:end] = (0.12 * env * (sig + noise)).astype(np.float32)
        pos = end + int(rng.uniform(0.1, 0.5) * sample_rate)
        seg += 1
    return np.clip(out, -1, 1)


EMB_SHAPE = dict(sample_rate=16000, num_mel_bins=80, m_channels=32, embed_dim=256, blocks=(3, 4, 6, 3))


def embedding_state_dict(seed: int = 0, shape: Dict = EMB_SHAPE) -> Dict[str, np.ndarray]:
    """WeSpeaker ResNet34 weights: Kaiming-normal convolutions, BatchNorm statistics / affine terms away from the
    identity so that the folding is exercised."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, np.ndarray] = {}

    def conv(name, cout, cin, k):
        std = math.sqrt(2.0 / (cin * k * k))
        sd[name] = rng.normal(0, std, (cout, cin, k, k)).astype(np.float32)

    def bn(name, c):
        sd[name + ".weight"] = rng.uniform(0.6, 1.2, c).astype(np.float32)
        sd[name + ".bias"] = rng.uniform(-0.1, 0.1, c).astype(np.float32)
        sd[name + ".running_mean"] = rng.uniform(-0.1, 0.1, c).astype(np.float32)
        sd[name + ".running_var"] = rng.uniform(0.6, 1.4, c).astype(np.float32)

    c0 = shape["m_channels"]
    conv("resnet.conv1.weight", c0, 1, 3)
    bn("resnet.bn1", c0)
    cin = c0
    for li, nb in enumerate(shape["blocks"]):
        cout = c0 << li
        for bi in range(nb):
            p = f"resnet.layer{li + 1}.{bi}"
            stride = 2 if (li > 0 and bi == 0) else 1
            conv(p + ".conv1.weight", cout, cin, 3)
            bn(p + ".bn1", cout)
            conv(p + ".conv2.weight", cout, cout, 3)
            bn(p + ".bn2", cout)
            sd[p + ".bn2.weight"] *= 0.5          # keeps the residual stream from growing over 16 blocks
            if stride != 1 or cin != cout:
                conv(p + ".shortcut.0.weight", cout, cin, 1)
                bn(p + ".shortcut.1", cout)
            cin = cout
    fq = shape["num_mel_bins"]
    for _ in range(3):
        fq = (fq - 1) // 2 + 1
    d = 2 * cin * fq
    bound = 1.0 / math.sqrt(d)
    sd["resnet.seg_1.weight"] = _uniform(rng, (shape["embed_dim"], d), bound)
    sd["resnet.seg_1.bias"] = _uniform(rng, (shape["embed_dim"],), bound)
    return sd
