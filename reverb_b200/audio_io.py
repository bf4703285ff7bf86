"""Audio file reading with the contract of the reference's `torchaudio.load(audio_file, normalize=False)`
(asr/wenet/cli/reverb.py:122): a (channels, frames) array + the sample rate, where INTEGER PCM WAV keeps its integer
sample values (the fbank front end works on int16-VALUED samples, cli/reverb.py:124, dataset/processor.py:361) and
everything else comes back as float32 in [-1, 1].

RIFF/WAVE is parsed here (PCM 8 / 16 / 24 / 32 bit, IEEE float 32 / 64, WAVE_FORMAT_EXTENSIBLE, any channel count) —
the stdlib `wave` module only does plain PCM.  Value conventions follow torchaudio: uint8 stays unsigned 0..255,
24-bit samples are left-justified in int32 (x << 8), 32-bit stay int32, float64 is narrowed to float32.
Any other container (flac, mp3, ogg, ...) is handed to torchaudio itself when it has a working decoder backend
(torchcodec / ffmpeg), which is exactly what the reference relies on; without one a clear error is raised.
"""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np

WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT, WAVE_FORMAT_EXTENSIBLE = 0x0001, 0x0003, 0xFFFE


def _parse_riff_wave(data: bytes, path: str) -> Tuple[np.ndarray, int]:
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            if len(body) < 16:
                raise ValueError(f"{path}: truncated fmt chunk")
            tag, nch, rate, _brate, block, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == WAVE_FORMAT_EXTENSIBLE and len(body) >= 40:
                tag = struct.unpack_from("<H", body, 24)[0]          # first two bytes of the SubFormat GUID
            fmt = (tag, nch, rate, block, bits)
        elif cid == b"data":
            pcm = body                                               # a streamed file may state size 0xFFFFFFFF: slice clips
        pos += 8 + size + (size & 1)                                 # chunks are word aligned
    if fmt is None or pcm is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, nch, rate, block, bits = fmt
    if nch < 1:
        raise ValueError(f"{path}: bad channel count {nch}")
    if tag not in (WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT):
        raise ValueError(f"{path}: unsupported WAVE format tag 0x{tag:04x} (only PCM and IEEE float)")
    if bits % 8 != 0 or bits == 0:
        raise ValueError(f"{path}: unsupported sample width {bits}")
    width = bits // 8
    nfr = len(pcm) // (width * nch)
    raw = pcm[:nfr * width * nch]
    if tag == WAVE_FORMAT_PCM:
        if bits == 8:
            x = np.frombuffer(raw, dtype=np.uint8)
        elif bits == 16:
            x = np.frombuffer(raw, dtype="<i2")
        elif bits == 24:
            b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            x = (b[:, 0] << 8) | (b[:, 1] << 16) | (b[:, 2] << 24)   # left-justified in int32, like torchaudio / libsndfile
            x = x.astype(np.int32)
        elif bits == 32:
            x = np.frombuffer(raw, dtype="<i4")
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    else:
        if bits == 32:
            x = np.frombuffer(raw, dtype="<f4")
        elif bits == 64:
            x = np.frombuffer(raw, dtype="<f8").astype(np.float32)
        else:
            raise ValueError(f"{path}: unsupported float width {bits}")
    return np.ascontiguousarray(x.reshape(nfr, nch).T), int(rate)


def load_audio(path: str) -> Tuple[np.ndarray, int]:
    """(channels, frames) samples + sample rate, `torchaudio.load(path, normalize=False)` semantics."""
    with open(path, "rb") as f:
        head = f.read(12)
        if head[:4] == b"RIFF" and head[8:12] == b"WAVE":
            return _parse_riff_wave(head + f.read(), str(path))
    try:                                       # any other container: the reference's own route, if a backend exists here
        import torchaudio
        wav, rate = torchaudio.load(str(path), normalize=False)
        return np.ascontiguousarray(wav.numpy()), int(rate)
    except Exception as e:
        raise ValueError(f"{path}: only RIFF/WAVE is decoded natively and torchaudio has no working decoder backend "
                         f"here for this file ({type(e).__name__}: {e}); convert it to WAV (e.g. `ffmpeg -i in out.wav`)") from e
