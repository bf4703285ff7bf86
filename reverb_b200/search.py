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

    l2r[i, j] = log p(w_j | w_<j) for j < U_i, l2r[i, U_i] = log p(eos), zeros behind; r2l likewise for the
    right-to-left decoder (already re-indexed to hypothesis order).  Float semantics follow the reference: the
    decoder scores are summed sequentially in float32 (0-d float32 tensors there) — `np.add.accumulate` is the same
    left-to-right float32 recurrence, and the trailing zeros are exact no-ops — while confidences and the CTC term
    go through double precision Python floats.
    """
    n = len(hyps)
    lens = np.fromiter((len(h) for h in hyps), dtype=np.int64, count=n)
    score = np.add.accumulate(np.ascontiguousarray(l2r[:n], dtype=np.float32), axis=1, dtype=np.float32)[:, -1]
    use_r = reverse_weight > 0 and r2l is not None
    if use_r:
        r_score = np.add.accumulate(np.ascontiguousarray(r2l[:n], dtype=np.float32), axis=1, dtype=np.float32)[:, -1]
        score = (score * np.float32(1 - reverse_weight) + r_score * np.float32(reverse_weight)).astype(np.float32)
    norm = (score / (lens + 1).astype(np.float32)).astype(np.float32)
    ctc_term = np.asarray([c * ctc_weight for c in ctc_scores[:n]], dtype=np.float64).astype(np.float32)
    total = (score + ctc_term).astype(np.float32)
    best = int(np.argmax(total)) if n and not np.isnan(total).any() else 0     # first maximum == strict `>` scan
    if n and np.isnan(total).any():                                            # keep the reference's scan semantics
        best, best_score = 0, -float("inf")
        for i in range(n):
            if float(total[i]) > best_score:
                best, best_score = i, float(total[i])
    U = int(lens[best])
    tc = [math.exp(float(l2r[best, j])) for j in range(U)]
    if use_r:
        tc = [(tc[j] + math.exp(float(r2l[best, j]))) / 2 for j in range(U)]
    return DecodeResult(hyps[best], float(total[best]), confidence=math.exp(float(norm[best])),
                        times=nbest_times[best], tokens_confidence=tc)


def rescoring_pick_batch(toks: np.ndarray, tims: np.ndarray, olen: np.ndarray, ctc_scores: np.ndarray,
                         nhyp: np.ndarray, l2r: np.ndarray, r2l: Optional[np.ndarray], ctc_weight: float,
                         reverse_weight: float) -> List[DecodeResult]:
    """`rescoring_pick` for a whole batch straight from the prefix-beam arrays (no per-hypothesis Python objects):
    toks/tims (B, N, L), olen (B, N, 2), ctc_scores (B, N) float64, nhyp (B,), l2r/r2l (B, N, Lmax+1) float32.
    Same float semantics as `rescoring_pick` (sequential float32 sums, see there)."""
    B, N = ctc_scores.shape
    lens = olen[:, :, 0].astype(np.int64)
    score = np.add.accumulate(l2r, axis=2, dtype=np.float32)[:, :, -1]
    use_r = reverse_weight > 0 and r2l is not None
    if use_r:
        r_score = np.add.accumulate(r2l, axis=2, dtype=np.float32)[:, :, -1]
        score = (score * np.float32(1 - reverse_weight) + r_score * np.float32(reverse_weight)).astype(np.float32)
    norm = (score / (lens + 1).astype(np.float32)).astype(np.float32)
    total = (score + (ctc_scores * ctc_weight).astype(np.float32)).astype(np.float32)
    valid = np.arange(N)[None, :] < nhyp[:, None]
    total = np.where(valid, total, -np.inf).astype(np.float32)
    best = np.argmax(total, axis=1)                      # first maximum == the reference's strict `>` scan
    out = []
    for b in range(B):
        i = int(best[b])
        U = int(lens[b, i])
        # exp in float64 of the float32 log-probs (numpy's vectorised exp; the reference's math.exp may differ by 1 ulp)
        tc = np.exp(l2r[b, i, :U].astype(np.float64))
        if use_r:
            tc = (tc + np.exp(r2l[b, i, :U].astype(np.float64))) / 2
        nt = int(olen[b, i, 1])
        out.append(DecodeResult(tuple(toks[b, i, :U].tolist()), float(total[b, i]),
                                confidence=math.exp(float(norm[b, i])), times=tims[b, i, :nt].tolist(),
                                tokens_confidence=tc.tolist()))
    return out


def attention_beam_search(step_topk, batch_size: int, maxlen: int, beam_size: int, sos: int, eos: int,
                          length_penalty: float = 0.0) -> List[DecodeResult]:
    """Host bookkeeping of the `attention` decode mode (transformer/search.py:251-360, non-whisper branch), numpy
    float32 like the reference's tensors.  step_topk(hyps (B*N, i) int64) -> (logp (B*N, N), index (B*N, N)) is the
    decoder step (Engine.decoder_step_topk).  Finished beams keep one zero-cost <eos> branch (utils/mask.py:257-303);
    the best beam per utterance is chosen after the length penalty; returns DecodeResult(tokens) only (no times /
    confidences: `transcribe(mode="attention")` fails in the reference for that reason)."""
    B, N = batch_size, beam_size
    running = B * N
    neg_inf = np.float32(-np.inf)
    hyps = np.full((running, 1), sos, dtype=np.int64)
    scores = np.tile(np.array([0.0] + [-np.inf] * (N - 1), dtype=np.float32), B).reshape(-1, 1)
    end_flag = np.zeros((running, 1), dtype=bool)
    zeros = np.zeros((running, 1), dtype=bool)
    for _ in range(1, maxlen + 1):
        if int(end_flag.sum()) == running:
            break
        logp, index = step_topk(hyps)
        logp = np.array(logp, dtype=np.float32, copy=True).reshape(running, N)
        index = np.array(index, dtype=np.int64, copy=True).reshape(running, N)
        if N > 1:
            unfinished = np.concatenate([zeros, np.repeat(end_flag, N - 1, axis=1)], axis=1)
            finished = np.concatenate([end_flag, np.repeat(zeros, N - 1, axis=1)], axis=1)
        else:
            unfinished, finished = zeros, end_flag
        logp[unfinished] = neg_inf
        logp[finished] = 0.0
        index[np.repeat(end_flag, N, axis=1)] = eos
        cand = (scores + logp).astype(np.float32).reshape(B, N * N)
        order = np.argsort(-cand, axis=1, kind="stable")[:, :N]            # topk, sorted, ties -> lowest index
        scores = np.take_along_axis(cand, order, axis=1).reshape(-1, 1)
        best_k_index = (np.arange(B)[:, None] * N * N + order).reshape(-1)
        best_k_pred = index.reshape(-1)[best_k_index]
        hyps = np.concatenate([hyps[best_k_index // N], best_k_pred[:, None]], axis=1)
        end_flag = hyps[:, -1:] == eos
    final = scores.reshape(B, N)
    lengths = (hyps != eos).sum(axis=1).reshape(B, N).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        final = (final / np.power(lengths, np.float32(length_penalty))).astype(np.float32)
    best = np.argmax(final, axis=1)
    out = []
    for b in range(B):
        hyp = hyps[b * N + int(best[b]), 1:]
        out.append(DecodeResult(hyp[hyp != eos].tolist()))
    return out
