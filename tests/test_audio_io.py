"""reverb_b200/audio_io.py against hand-built RIFF/WAVE files: torchaudio.load(normalize=False) value conventions
(asr/wenet/cli/reverb.py:122) for every sample format the parser accepts."""
import struct
import wave

import numpy as np
import pytest

from reverb_b200.audio_io import load_audio


def _riff(fmt_body: bytes, data: bytes, extra: bytes = b"") -> bytes:
    chunks = b"fmt " + struct.pack("<I", len(fmt_body)) + fmt_body + extra + b"data" + struct.pack("<I", len(data)) + data
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


def _fmt(tag, nch, rate, bits):
    block = nch * bits // 8
    return struct.pack("<HHIIHH", tag, nch, rate, rate * block, block, bits)


def test_pcm16_matches_the_stdlib_reader(tmp_path):
    x = (np.arange(-300, 300) * 100).astype(np.int16).reshape(-1, 2)
    p = tmp_path / "a.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(8000); w.writeframes(x.tobytes())
    got, rate = load_audio(str(p))
    assert rate == 8000 and got.dtype == np.int16 and got.shape == (2, 300)
    assert np.array_equal(got, x.T)


def test_other_widths_and_float_and_extensible(tmp_path):
    rate = 44100
    # 8-bit unsigned stays 0..255
    u8 = np.array([0, 128, 255, 7], dtype=np.uint8)
    (tmp_path / "u8.wav").write_bytes(_riff(_fmt(1, 1, rate, 8), u8.tobytes()))
    got, r = load_audio(str(tmp_path / "u8.wav"))
    assert r == rate and got.dtype == np.uint8 and np.array_equal(got[0], u8)
    # 24-bit: left-justified in int32
    vals = np.array([0, 1, -1, 8388607, -8388608, 123456], dtype=np.int64)
    raw = b"".join(int(v & 0xFFFFFF).to_bytes(3, "little") for v in vals)
    (tmp_path / "s24.wav").write_bytes(_riff(_fmt(1, 1, rate, 24), raw, extra=b"LIST" + struct.pack("<I", 3) + b"abc\x00"))
    got, _ = load_audio(str(tmp_path / "s24.wav"))
    assert got.dtype == np.int32 and np.array_equal(got[0].astype(np.int64), vals * 256)
    # 32-bit int, 3 channels
    x32 = np.arange(-6, 6, dtype=np.int32).reshape(4, 3) * 1000003
    (tmp_path / "s32.wav").write_bytes(_riff(_fmt(1, 3, rate, 32), x32.tobytes()))
    got, _ = load_audio(str(tmp_path / "s32.wav"))
    assert got.shape == (3, 4) and np.array_equal(got, x32.T)
    # float32 / float64 -> float32
    f = np.linspace(-1, 1, 10, dtype=np.float32)
    (tmp_path / "f32.wav").write_bytes(_riff(_fmt(3, 1, rate, 32), f.tobytes()))
    (tmp_path / "f64.wav").write_bytes(_riff(_fmt(3, 1, rate, 64), f.astype(np.float64).tobytes()))
    for name in ("f32.wav", "f64.wav"):
        got, _ = load_audio(str(tmp_path / name))
        assert got.dtype == np.float32 and np.array_equal(got[0], f)
    # WAVE_FORMAT_EXTENSIBLE wrapping 16-bit PCM
    x16 = np.array([1, -2, 3, -4], dtype=np.int16)
    ext = _fmt(0xFFFE, 1, rate, 16) + struct.pack("<HHI", 22, 16, 4) + struct.pack("<H", 1) + b"\x00" * 14
    (tmp_path / "ext.wav").write_bytes(_riff(ext, x16.tobytes()))
    got, _ = load_audio(str(tmp_path / "ext.wav"))
    assert got.dtype == np.int16 and np.array_equal(got[0], x16)


def test_errors_are_loud(tmp_path):
    (tmp_path / "bad.flac").write_bytes(b"fLaC" + b"\x00" * 64)
    with pytest.raises(ValueError, match="WAV"):
        load_audio(str(tmp_path / "bad.flac"))
    (tmp_path / "adpcm.wav").write_bytes(_riff(_fmt(2, 1, 8000, 4), b"\x00" * 16))
    with pytest.raises(ValueError, match="format tag"):
        load_audio(str(tmp_path / "adpcm.wav"))
