"""Chunk-sharded long-form decoding over the GPUs of one node (one process per GPU).

Chunks of a recording never exchange state in the reference (asr/wenet/cli/reverb.py:214-234: every
`decode` call sees only its own (B, T, 80) slice; the only cross-chunk quantity is the host-side
`time_shift_ms`, :314-319), so the path shards as independent units:

  * contiguous block partition of the chunk index range over the ranks (rank order == chunk order, so
    the gather needs no permutation);
  * each rank computes fbank for its own sample range — frame i depends only on samples
    [160 i, 160 i + 400) — and runs fbank -> encoder -> searches -> rescoring on its chunks;
  * ONE collective per file: an all-gather of fixed-size per-chunk result records (NCCL over NVLink in
    production; gloo in the CPU tests).  No collective sits on the data path.
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch

from .search import DecodeResult

FRAME_SHIFT, FRAME_LEN = 160, 400


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: rank r owns [r*ceil(n/W), min(n, (r+1)*ceil(n/W)))."""
    per = -(-n_units // world) if n_units > 0 else 0
    lo = min(n_units, rank * per)
    return lo, min(n_units, lo + per)


def num_frames(n_samples: int) -> int:
    return 0 if n_samples < FRAME_LEN else 1 + (n_samples - FRAME_LEN) // FRAME_SHIFT


def chunk_plan(n_samples: int, chunk_size: int) -> Tuple[int, int]:
    """(total fbank frames, number of chunks) of a recording cut into `chunk_size`-frame chunks."""
    m = num_frames(n_samples)
    return m, (math.ceil(m / chunk_size) if m > 0 else 0)


def sample_range_for_chunks(c0: int, c1: int, chunk_size: int, total_frames: int) -> Tuple[int, int, int]:
    """Samples [s0, s1) a rank must read to compute the frames of chunks [c0, c1), and that frame count.
    Neighbouring ranks overlap by 240 samples (window 400, shift 160); nothing else couples them."""
    f0 = c0 * chunk_size
    f1 = min(total_frames, c1 * chunk_size)
    if f1 <= f0:
        return 0, 0, 0
    return FRAME_SHIFT * f0, FRAME_SHIFT * (f1 - 1) + FRAME_LEN, f1 - f0


# ------------------------------------------------------------------------------------------------------
# fixed-size result records: int32 words
#   [0] n_tokens (-1 = empty slot)  [1] n_times (-1 = None)  [2] has_conf  [3:5] score (float64 bits)
#   [5:7] confidence (float64 bits)  then tokens[U], times[U], token_conf[U] (float32 bits)
def record_words(max_tokens: int) -> int:
    return 7 + 3 * max_tokens


def pack_results(results: Sequence[DecodeResult], n_slots: int, max_tokens: int) -> np.ndarray:
    rec = np.zeros((n_slots, record_words(max_tokens)), dtype=np.int32)
    rec[:, 0] = -1
    for i, r in enumerate(results):
        toks = list(r.tokens)
        if len(toks) > max_tokens:
            raise ValueError(f"hypothesis of {len(toks)} tokens does not fit a record of {max_tokens}")
        rec[i, 0] = len(toks)
        rec[i, 1] = -1 if r.times is None else len(r.times)
        rec[i, 2] = 0 if r.tokens_confidence is None else 1
        rec[i, 3:5] = np.array([float(r.score)], dtype=np.float64).view(np.int32)
        rec[i, 5:7] = np.array([float(r.confidence)], dtype=np.float64).view(np.int32)
        U = max_tokens
        rec[i, 7:7 + len(toks)] = toks
        if r.times is not None:
            rec[i, 7 + U:7 + U + len(r.times)] = r.times
        if r.tokens_confidence is not None:
            rec[i, 7 + 2 * U:7 + 2 * U + len(toks)] = np.asarray(r.tokens_confidence, dtype=np.float32).view(np.int32)
    return rec


def unpack_results(rec: np.ndarray, max_tokens: int) -> List[DecodeResult]:
    out = []
    U = max_tokens
    for row in rec:
        n = int(row[0])
        if n < 0:
            continue
        score = float(row[3:5].copy().view(np.float64)[0])
        conf = float(row[5:7].copy().view(np.float64)[0])
        times = None if row[1] < 0 else row[7 + U:7 + U + int(row[1])].tolist()
        tc = None
        if row[2]:
            tc = [float(x) for x in row[7 + 2 * U:7 + 2 * U + n].copy().view(np.float32)]
        out.append(DecodeResult(row[7:7 + n].tolist(), score, conf, tc, times))
    return out


def gather_records(rec: np.ndarray, device: torch.device) -> np.ndarray:
    """all-gather of every rank's (per, W) int32 record block -> (world * per, W), rank order == chunk order."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return rec
    if device.type == "cuda":
        src = torch.from_numpy(rec).pin_memory().to(device, non_blocking=True)
    else:
        src = torch.from_numpy(rec)
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=torch.int32, device=src.device)
    dist.all_gather_into_tensor(out, src.contiguous())
    return out.cpu().numpy()


def gather_chunk_results(local: Sequence[DecodeResult], n_chunks: int, max_tokens: int,
                         device: torch.device) -> List[DecodeResult]:
    """The single collective of the path: all-gather of the per-chunk records, in chunk order."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    per = -(-n_chunks // world) if n_chunks > 0 else 0
    return unpack_results(gather_records(pack_results(local, per, max_tokens), device), max_tokens)


class RecordGatherer:
    """The same all-gather, asynchronous: packs on the host, then H2D -> NCCL all-gather -> D2H on a SIDE stream, so a
    pipelined decoder (ASRModel.decode_stream) keeps enqueueing the next batches on the main stream; `wait` returns
    the gathered records.  One instance per fixed (chunks per rank, max_tokens) shape."""

    def __init__(self, device: torch.device, per_rank: int, max_tokens: int):
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.device, self.per, self.max_tokens = device, per_rank, max_tokens
        self.side = torch.cuda.Stream(device=device) if device.type == "cuda" else None

    def submit(self, local: Sequence[DecodeResult]):
        rec = pack_results(local, self.per, self.max_tokens)
        if self.world == 1:
            return {"host": torch.from_numpy(rec), "ev": None}
        if self.side is None:                       # CPU process group (gloo, tests): synchronous
            return {"host": torch.from_numpy(gather_records(rec, self.device)), "ev": None}
        src_h = torch.from_numpy(rec).pin_memory()
        out_h = torch.empty((self.world * self.per, rec.shape[1]), dtype=torch.int32, pin_memory=True)
        with torch.cuda.stream(self.side):
            src = src_h.to(self.device, non_blocking=True)
            out = torch.empty(out_h.shape, dtype=torch.int32, device=self.device)
            self.dist.all_gather_into_tensor(out, src)
            out_h.copy_(out, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.side)
        return {"host": out_h, "ev": ev, "keep": (src_h, src, out)}

    def wait(self, h) -> np.ndarray:
        if h["ev"] is not None:
            h["ev"].synchronize()
        return h["host"].numpy()


def decode_sharded(decode_chunks: Callable[[int, int], List[DecodeResult]], n_chunks: int, max_tokens: int,
                   device: torch.device) -> List[DecodeResult]:
    """Run `decode_chunks(c0, c1)` on this rank's block of chunks, then gather everybody's results."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    c0, c1 = shard_range(n_chunks, rank, world)
    local = decode_chunks(c0, c1) if c1 > c0 else []
    assert len(local) == c1 - c0
    return gather_chunk_results(local, n_chunks, max_tokens, device)


def transcribe_sharded(asr, pcm: np.ndarray, mode: str = "attention_rescoring", chunk_size: int = 2051,
                       batch_size: int = 8, beam_size: int = 10, ctc_weight: float = 0.1,
                       reverse_weight: float = 0.0, verbatimicity: float = 1.0, blank_penalty: float = 0.0,
                       max_tokens: int = 0) -> List[DecodeResult]:
    """Long-form decode of int16 samples `pcm` (16 kHz mono), sharded over the initialised process group.
    Returns the per-chunk DecodeResults of the WHOLE file on every rank (feed them to reverb.get_output)."""
    total_frames, n_chunks = chunk_plan(int(pcm.shape[0]), chunk_size)
    max_tokens = max_tokens or asr.engine.encoder_out_frames(chunk_size)
    cat = torch.tensor([verbatimicity, 1.0 - verbatimicity])

    def decode_chunks(c0: int, c1: int) -> List[DecodeResult]:
        s0, s1, nfr = sample_range_for_chunks(c0, c1, chunk_size, total_frames)
        wave = torch.from_numpy(np.ascontiguousarray(pcm[s0:s1])).pin_memory().to(asr.device, non_blocking=True)
        feats = asr.engine.fbank(wave)[:nfr].unsqueeze(0)
        out: List[DecodeResult] = []
        for res in asr.model.decode_stream(asr.feats_batcher(feats, chunk_size, batch_size), [mode], beam_size,
                                           ctc_weight=ctc_weight, reverse_weight=reverse_weight, blank_id=asr.blank_id,
                                           blank_penalty=blank_penalty, cat_embs=cat):
            out.extend(res[mode])
        return out

    return decode_sharded(decode_chunks, n_chunks, max_tokens, asr.device)
