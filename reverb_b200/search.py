"""Host side of the decoding methods: result container and the (tiny) score arithmetic that
stays on the CPU.  The searches themselves run on the GPU (csrc/ctc.cu) through Engine.

Mirrors asr/wenet/transformer/search.py of the reference: `DecodeResult` (:29-58), the
output contract of `ctc_greedy_search` (:106-121), `ctc_prefix_beam_search` (:124-248)
and the score combination of `attention_rescoring` (:413-447).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np


class DecodeResult:
    """Same fields, defaults and meaning as the reference's DecodeResult (search.py:29-58)."""

    def __init__(self, tokens, score: float = 0.0, confidence: float = 0.0,
                 tokens_confidence: Optional[List[float]] = None, times: Optional[List[int]] = None,
                 nbest=None, nbest_scores: Optional[List[float]] = None, nbest_times=None):
        self.tokens = tokens
        self.score = score
        self.confidence = confidence
        self.tokens_confidence = tokens_confidence
        self.times = times
        self.nbest = nbest
        self.nbest_scores = nbest_scores
        self.nbest_times = nbest_times

    def __repr__(self):
        return f"DecodeResult(tokens={list(self.tokens)}, score={self.score}, confidence={self.confidence})"


def greedy_results(token_lists: Sequence[Sequence[int]]) -> List[DecodeResult]:
    # the reference's greedy search returns tokens only (times / confidence stay None, search.py:119)
    return [DecodeResult(list(t)) for t in token_lists]


def prefix_beam_results(per_utt) -> List[DecodeResult]:
    """per_utt: Engine.prefix_beam_search output.  Best hypothesis = first of the n-best (search.py:235-247)."""
    out = []
    for nbest, scores, times in per_utt:
        out.append(DecodeResult(tokens=nbest[0], score=scores[0], times=times[0], nbest=nbest,
                                nbest_scores=scores, nbest_times=times))
    return out


def rescoring_pick(hyps: Sequence[tuple], ctc_scores: Sequence[float], nbest_times, l2r: np.ndarray,
                   r2l: Optional[np.ndarray], ctc_weight: float, reverse_weight: float) -> DecodeResult:
    """Combine decoder and CTC scores and keep the first strict maximum (search.py:413-447).

    l2r[i, j] = log p(w_j | w_<j) for j < U_i and l2r[i, U_i] = log p(eos); r2l likewise for the
    right-to-left decoder (already re-indexed to hypothesis order).  Like the reference, the decoder
    scores are accumulated in float32 (they are 0-d float32 tensors there); confidences and the CTC
    term use double precision Python floats.
    """
    best_score, best_index = -float("inf"), 0
    confidences, tok_conf = [], []
    rw32 = np.float32(reverse_weight)
    for i, hyp in enumerate(hyps):
        U = len(hyp)
        score = np.float32(0.0)
        tc = []
        for j in range(U):
            s = np.float32(l2r[i, j])
            score = np.float32(score + s)
            tc.append(math.exp(float(s)))
        score = np.float32(score + np.float32(l2r[i, U]))
        if reverse_weight > 0 and r2l is not None:
            r_score = np.float32(0.0)
            for j in range(U):
                s = np.float32(r2l[i, j])
                r_score = np.float32(r_score + s)
                tc[j] = (tc[j] + math.exp(float(s))) / 2
            r_score = np.float32(r_score + np.float32(r2l[i, U]))
            score = np.float32(np.float32(score * np.float32(1 - reverse_weight)) + np.float32(r_score * rw32))
        confidences.append(math.exp(float(np.float32(score / np.float32(U + 1)))))
        score = np.float32(score + np.float32(ctc_scores[i] * ctc_weight))
        if float(score) > best_score:
            best_score, best_index = float(score), i
        tok_conf.append(tc)
    return DecodeResult(hyps[best_index], best_score, confidence=confidences[best_index],
                        times=nbest_times[best_index], tokens_confidence=tok_conf[best_index])
