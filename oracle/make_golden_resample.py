"""ORACLE tooling: golden vectors for the resampling front-end straight from torchaudio.transforms.Resample (the
reference's call, asr/wenet/cli/reverb.py:125-128).  Run from the repo root: python oracle/make_golden_resample.py"""
import os
import sys

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from reverb_b200 import synth
    out = {}
    for rate, n in [(8000, 2999), (22050, 8001), (44100, 16000), (48000, 17357), (11025, 401)]:
        pcm = synth.synth_audio(n / 16000.0 + 1e-9, seed=200 + rate % 97)[:n]     # n int16-valued samples "at `rate`"
        assert pcm.shape[0] == n
        w = torch.from_numpy(pcm.astype(np.float32)).unsqueeze(0)
        y = torchaudio.transforms.Resample(orig_freq=rate, new_freq=16000)(w)
        out[f"r{rate}_n{n}_seed{200 + rate % 97}"] = y[0].numpy()
        print(rate, n, "->", y.shape[1])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resample.npz"), **out)


if __name__ == "__main__":
    main()
