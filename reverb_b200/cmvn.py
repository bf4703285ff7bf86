"""Global CMVN statistics loader (mean, inverse std) — asr/wenet/utils/cmvn.py:21-93."""
from __future__ import annotations

import json
import math

import numpy as np


def _finish(sums, sq_sums, count):
    mean = np.asarray(sums, dtype=np.float64) / count
    var = np.asarray(sq_sums, dtype=np.float64) / count - mean * mean
    var = np.maximum(var, 1.0e-20)
    return mean, 1.0 / np.sqrt(var)


def load_cmvn(path: str, is_json: bool):
    if is_json:
        with open(path) as f:
            st = json.load(f)
        return _finish(st["mean_stat"], st["var_stat"], st["frame_num"])
    # kaldi text matrix from `compute-cmvn-stats --binary=false`: "[ sum_1..sum_D count sq_1..sq_D 0 ]"
    with open(path) as f:
        txt = f.read()
    if txt[:2] == "\0B":
        raise ValueError("kaldi binary cmvn is not supported, recompute with --binary=false")
    arr = txt.split()
    if not (arr and arr[0] == "[" and arr[-1] == "]" and arr[-2] == "0"):
        raise ValueError(f"{path}: not a kaldi text cmvn file")
    dim = (len(arr) - 4) // 2
    sums = [float(x) for x in arr[1:1 + dim]]
    count = float(arr[1 + dim])
    sq = [float(x) for x in arr[2 + dim:2 + 2 * dim]]
    return _finish(sums, sq, count)
