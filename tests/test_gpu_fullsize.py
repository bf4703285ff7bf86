"""Parity at BASELINE.json's FULL sizes (reverb_asr_v1 shape d=1024, L=18, V=10001; 64 x 30 s chunks per batch)
through size-independent properties — the CPU oracle needs minutes per chunk at this shape, so instead of
recomputing it these tests pin what must hold whatever the weights are:

  * fbank is frame-local: shifting the signal by k hops shifts the features by k frames, bit for bit
    (torchaudio kaldi.fbank with dither=0, subtract_mean=False; SURVEY.md §8e uses this to shard a file)
  * GEMM: power-of-two scaling commutes with every rounding step (out(2A) == 2 out(A) bit-exact) and rows are
    independent (out(A[perm]) == out(A)[perm]) at the FFN shape M=47872, N=4096, K=1024
  * chunks are independent units (reverb.py:214-234): decoding is deterministic, equivariant under a permutation of
    the batch, and a sub-batch decodes to exactly what it decodes to inside the full batch
  * a zero-padded tail chunk of a causal model gives the valid frames the same encoder output as the unpadded
    chunk
  * structural invariants of the search outputs (search.py:124-248, 363-448): n-best sorted by score, unique,
    one peak time per token, non-decreasing and inside the utterance; rescoring picks a member of the n-best.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CHUNK_FRAMES = 2998
CHUNK_SAMPLES = 480000
B_FULL = 64


@pytest.fixture(scope="module")
def big(bench_model_dir):
    import reverb_b200
    from reverb_b200 import synth
    d = bench_model_dir
    asr = reverb_b200.ReverbASR(os.path.join(d, "config.yaml"), os.path.join(d, "synth.pt"), gpu=0)
    base = [synth.synth_audio(30.0, seed=4321 + i) for i in range(4)]
    pcm = np.empty((B_FULL, CHUNK_SAMPLES), dtype=np.int16)
    for i in range(B_FULL):
        pcm[i] = (base[i % 4].astype(np.float32) * (1.0 - 0.04 * (i // 4 % 16))).astype(np.int16)
    feats = asr.engine.fbank_batch(torch.from_numpy(pcm).cuda())
    assert feats.shape == (B_FULL, CHUNK_FRAMES, 80)
    return asr, pcm, feats


def _decode(asr, feats, lens, modes=("ctc_prefix_beam_search", "attention_rescoring"), rw=0.0):
    cat = torch.tensor([1.0, 0.0])
    return asr.model.decode(list(modes), feats, lens, 10, ctc_weight=0.1, reverse_weight=rw,
                            blank_id=asr.blank_id, cat_embs=cat)


def _same(a, b):
    assert tuple(a.tokens) == tuple(b.tokens) and a.times == b.times
    assert float(a.score) == float(b.score)
    if a.nbest is not None:
        assert [tuple(h) for h in a.nbest] == [tuple(h) for h in b.nbest]
        assert a.nbest_times == b.nbest_times and list(a.nbest_scores) == list(b.nbest_scores)
    if a.tokens_confidence is not None:
        assert list(a.tokens_confidence) == list(b.tokens_confidence) and a.confidence == b.confidence


def test_fbank_is_frame_local_at_full_length(big):
    asr, pcm, feats = big
    k = 37
    x = torch.from_numpy(pcm[:4]).cuda()
    shifted = asr.engine.fbank_batch(x[:, 160 * k:].contiguous())
    assert torch.equal(shifted, feats[:4, k:k + shifted.shape[1]])
    assert shifted.shape[1] == CHUNK_FRAMES - k


def test_gemm_scaling_and_row_independence_at_ffn_shape():
    from reverb_b200 import _lib
    lib = _lib.load()
    M, N, K = 47872, 4096, 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())

    def run(a, b, act):
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        assert lib.rvb_gemm_bf16(p(a), p(W), p(b) if b is not None else None, M, N, K, act, 0, 1.0, p(out), N, st) == 0
        return out
    base = run(A, None, 0)
    assert torch.equal(run((A.float() * 2).bfloat16(), None, 0).float(), base.float() * 2)
    perm = torch.randperm(M, device="cuda", generator=g)
    silu = run(A, bias, 2)
    assert torch.equal(run(A[perm].contiguous(), bias, 2), silu[perm])
    # spot-check 64 rows against fp32 math
    rows = perm[:64]
    ref = torch.nn.functional.silu(A[rows].float() @ W.float().t() + bias)
    torch.testing.assert_close(silu[rows].float(), ref, rtol=2e-2, atol=2e-2)


def test_decode_is_deterministic_permutation_equivariant_and_batch_invariant(big):
    asr, pcm, feats = big
    lens = torch.full((B_FULL,), CHUNK_FRAMES, dtype=torch.int32)
    r1 = _decode(asr, feats, lens)
    r2 = _decode(asr, feats, lens)
    perm = torch.from_numpy(np.random.default_rng(0).permutation(B_FULL))
    rp = _decode(asr, feats[perm.cuda()].contiguous(), lens)
    sub = _decode(asr, feats[:8].contiguous(), lens[:8])
    for mode in ("ctc_prefix_beam_search", "attention_rescoring"):
        for b in range(B_FULL):
            _same(r1[mode][b], r2[mode][b])
            _same(rp[mode][b], r1[mode][int(perm[b])])
    for b in range(8):
        _same(sub["ctc_prefix_beam_search"][b], r1["ctc_prefix_beam_search"][b])
        # the decoder batch is padded to a different length, so compare the decision, not the float score
        assert tuple(sub["attention_rescoring"][b].tokens) == tuple(r1["attention_rescoring"][b].tokens)
        assert abs(float(sub["attention_rescoring"][b].score) - float(r1["attention_rescoring"][b].score)) < 1e-3


def test_search_output_invariants_at_full_size(big):
    asr, pcm, feats = big
    lens = torch.full((B_FULL,), CHUNK_FRAMES, dtype=torch.int32)
    for rw in (0.0, 0.3):
        res = _decode(asr, feats, lens, rw=rw)
        V = 10001
        for b in range(B_FULL):
            pb, ar = res["ctc_prefix_beam_search"][b], res["attention_rescoring"][b]
            assert 1 <= len(pb.nbest) <= 10 and len(set(map(tuple, pb.nbest))) == len(pb.nbest)
            assert all(s0 >= s1 for s0, s1 in zip(pb.nbest_scores, pb.nbest_scores[1:]))
            assert tuple(pb.tokens) == tuple(pb.nbest[0]) and pb.times == pb.nbest_times[0]
            for h, tm in zip(pb.nbest, pb.nbest_times):
                assert len(h) == len(tm) and all(0 < t < V for t in h)
                assert all(0 <= t < 748 for t in tm) and all(a <= c for a, c in zip(tm, tm[1:]))
            assert tuple(ar.tokens) in set(map(tuple, pb.nbest))
            i = [tuple(h) for h in pb.nbest].index(tuple(ar.tokens))
            assert ar.times == pb.nbest_times[i]
            assert 0.0 < ar.confidence <= 1.0 and len(ar.tokens_confidence) == len(ar.tokens)
            assert all(0.0 < c <= 1.0 + 1e-6 for c in ar.tokens_confidence)
            assert math.isfinite(float(ar.score))


def test_zero_padded_tail_chunk_equals_unpadded_chunk_for_the_causal_model(big):
    asr, pcm, feats = big
    valid = 998                                           # -> 248 encoder frames
    padded = feats[:4].clone()
    padded[:, valid:] = 0.0                               # the reference pads with 0.0 in feature space, before CMVN
    cat = torch.tensor([1.0, 0.0])
    enc_p, len_p = asr.model._forward_encoder(padded, torch.full((4,), valid, dtype=torch.int32), cat)
    enc_u, len_u = asr.model._forward_encoder(feats[:4, :valid].contiguous(), torch.full((4,), valid, dtype=torch.int32), cat)
    assert list(len_p) == list(len_u) == [248] * 4
    a, b = enc_p[:, :248].float(), enc_u[:, :248].float()
    rel = float(((a - b) ** 2).mean().sqrt() / (b ** 2).mean().sqrt())
    assert rel < 1e-6, rel
