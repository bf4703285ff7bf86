"""Host glue of the diarization pipeline (reverb_b200/diarization/pipeline.py) — no GPU: the two networks are replaced by
stubs that read a ground-truth speaker timeline, so what is tested is the windowing, aggregation, counting, clustering,
reconstruction and run-length encoding that the pipeline restates from pyannote's published algorithm
(** parity unpinned ** against pyannote itself)."""
import io

import numpy as np
import torch

from reverb_b200.diarization import pipeline as P
from reverb_b200.diarization.rttm import load_rttm


def test_receptive_field_and_chunking():
    rf = P.receptive_field()
    assert abs(rf.duration - 991 / 16000) < 1e-12 and abs(rf.step - 270 / 16000) < 1e-12
    assert rf.closest_frame(0.5 * rf.duration) == 0
    assert P.chunk_starts(160000, 160000, 16000) == (1, False)
    assert P.chunk_starts(160001, 160000, 16000) == (1, True)
    assert P.chunk_starts(100, 160000, 16000) == (0, True)
    assert P.chunk_starts(160000 + 5 * 16000, 160000, 16000) == (6, False)


def test_aggregate_averages_overlaps_and_skips_nan():
    chunks = P.SlidingWindow(0.0, 4.0, 2.0)
    frames = P.SlidingWindow(0.0, 1.0, 1.0)
    s = np.zeros((2, 4, 1), np.float32)
    s[0, :, 0] = [1, 1, 1, 1]
    s[1, :, 0] = [3, 3, np.nan, 3]
    avg = P.aggregate(s, chunks, frames)
    # window 1 starts at frame 2: frames 2,3 overlap -> (1+3)/2 ; frame 4 is NaN in the only window covering it
    assert avg.shape[0] >= 6
    assert np.allclose(avg[:4, 0], [1, 1, 2, 2])
    assert np.isnan(avg[4, 0]) and avg[5, 0] == 3
    tot = P.aggregate(s, chunks, frames, missing=0.0, skip_average=True)
    assert np.allclose(tot[:6, 0], [1, 1, 4, 4, 0, 3])


def test_agglomerative_clustering_two_blobs_and_small_cluster_merge():
    rng = np.random.default_rng(0)
    a = np.array([1.0, 0.0, 0.0]) + 0.05 * rng.normal(size=(30, 3))
    b = np.array([0.0, 1.0, 0.0]) + 0.05 * rng.normal(size=(30, 3))
    stray = np.array([[0.6, 0.0, 0.8]])                          # far from both: its own tiny cluster -> merged into A's
    x = np.vstack([a, b, stray])
    cl = P.agglomerative_clustering(x.copy(), threshold=0.7, min_cluster_size=5)
    assert len(set(cl[:30])) == 1 and len(set(cl[30:60])) == 1 and cl[0] != cl[30]
    assert cl[60] == cl[0] and set(cl) == {0, 1}
    assert P.agglomerative_clustering(x[:1].copy(), 0.7, 5).tolist() == [0]


def test_binarize_uses_frame_middles_and_merges_short_gaps():
    frames = P.SlidingWindow(0.0, 0.1, 0.1)
    act = np.zeros((20, 2), np.float32)
    act[2:6, 0] = 1
    act[7:10, 0] = 1
    act[15:, 1] = 1
    r = P.binarize(act, frames)
    assert [(round(a, 2), round(b, 2), k) for a, b, k in r] == [(0.25, 0.65, 0), (0.75, 1.05, 0), (1.55, 1.95, 1)]
    r = P.binarize(act, frames, min_duration_off=0.15)
    assert [(round(a, 2), round(b, 2), k) for a, b, k in r] == [(0.25, 1.05, 0), (1.55, 1.95, 1)]


class _Truth:
    """ground truth: speaker id per 10 ms (or -1), three speakers taking turns with one overlap region"""

    def __init__(self, seconds=40.0):
        n = int(seconds * 100)
        self.active = np.zeros((3, n), bool)
        self.active[0, 100:900] = True
        self.active[1, 850:1800] = True                            # overlaps speaker 0 for 0.5 s
        self.active[2, 2000:2900] = True
        self.active[0, 3000:3800] = True

    def at(self, t):
        i = np.clip((np.asarray(t) * 100).astype(int), 0, self.active.shape[1] - 1)
        return self.active[:, i]                                   # (3, len(t))


def _stub_networks(truth, rf, sample_rate=16000):
    protos = np.eye(3, 8, dtype=np.float32) * 4 + 0.5
    powerset = {(): 0, (0,): 1, (1,): 2, (2,): 3, (0, 1): 4, (0, 2): 5, (1, 2): 6}

    def seg(chunks):                                               # chunk c starts c seconds in (step 1 s): read it back
        B = chunks.shape[0]
        out = torch.full((B, 589, 7), -20.0)
        for b in range(B):
            t0 = float(chunks[b, 0])                               # windows carry their start time in sample 0
            times = t0 + np.array([rf.middle(i) for i in range(589)])
            act = truth.at(times)
            # local speaker order = order of first appearance inside the window (a permutation of the global ids)
            order = [s for s in np.argsort([np.argmax(a) if a.any() else 10**6 for a in act]) if act[s].any()]
            local = {g: i for i, g in enumerate(order)}
            for f in range(589):
                key = tuple(sorted(local[g] for g in range(3) if act[g, f]))[:2]
                out[b, f, powerset[key]] = 0.0
        return out

    def emb(chunks, weights):
        B, S, _ = weights.shape
        out = torch.zeros(B, S, 8)
        for b in range(B):
            t0 = float(chunks[b, 0])
            times = t0 + np.array([rf.middle(i) for i in range(589)])
            act = truth.at(times).astype(np.float32)               # (3, 589)
            for s in range(S):
                w = weights[b, s].numpy()
                score = act @ w
                g = int(np.argmax(score)) if score.max() > 0 else 0
                out[b, s] = torch.from_numpy(protos[g] + 0.01 * np.sin(np.arange(8) + b + s).astype(np.float32))
        return out

    return seg, emb


def test_pipeline_recovers_a_three_speaker_timeline_with_stub_networks():
    truth = _Truth(40.0)
    rf = P.receptive_field()
    seg, emb = _stub_networks(truth, rf)
    pipe = P.SpeakerDiarization(seg, emb, min_cluster_size=3, device="cpu")
    n = 40 * 16000
    wave = torch.zeros(n)
    # the stub networks find a window's position through its first sample: sample i of the recording holds i / sr
    wave[:] = torch.arange(n, dtype=torch.float32) / 16000.0
    turns = pipe.apply(wave)
    labels = sorted({t.label for t in turns})
    assert labels == ["SPEAKER_00", "SPEAKER_01", "SPEAKER_02"]
    # per-frame agreement with the truth (up to the label permutation given by first appearance)
    grid = np.arange(0, 40, 0.05)
    got = np.zeros((3, len(grid)), bool)
    for t in turns:
        got[int(t.label[-2:]), (grid >= t.start) & (grid < t.end)] = True
    want = truth.at(grid)
    best = max(np.mean(got[list(p)] == want) for p in [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)])
    assert best > 0.97, best
    # RTTM round trip
    buf = io.StringIO()
    pipe.write_rttm(buf, "rec", turns)
    buf.seek(0)
    back = load_rttm(buf)["rec"]
    assert len(back) == len(turns) and all(abs(a.start - b.start) < 1e-3 for a, b in zip(back, turns))


def test_cli_audio_reader_and_model_dir_errors(tmp_path):
    """`infer.read_audio`: 16 kHz int16 stereo WAV -> mono float32 in [-1, 1] (pyannote's Audio(mono='downmix'));
    `load_pipeline` without weights fails loudly instead of falling back to anything."""
    import wave
    import pytest
    from reverb_b200.diarization import infer
    path = tmp_path / "st.wav"
    left = (np.sin(np.arange(16000) * 0.05) * 16000).astype(np.int16)
    right = np.zeros(16000, np.int16)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(np.stack([left, right], axis=1).tobytes())
    x = infer.read_audio(str(path))
    assert x.dtype == np.float32 and x.shape == (16000,)
    assert np.allclose(x, left.astype(np.float32) / 32768.0 / 2, atol=1e-6)
    with pytest.raises(ValueError):
        infer.load_pipeline(str(tmp_path / "missing_dir"))


def test_gpu_pdist_route_gives_the_same_dendrogram_cut():
    """`condensed_euclidean` (torch, run on the CPU here) == scipy `pdist`, and clustering through it equals the direct
    scipy route."""
    from scipy.spatial.distance import pdist
    rng = np.random.default_rng(3)
    x = rng.normal(size=(400, 32))
    x[:, :4] += np.eye(3, 4)[rng.integers(0, 3, 400)] * 5
    xn = x / np.linalg.norm(x, axis=1, keepdims=True)
    assert np.allclose(P.condensed_euclidean(xn, "cpu"), pdist(xn), rtol=0, atol=1e-12)
    a = P.agglomerative_clustering(x.copy(), 0.9, 12)
    b = P.agglomerative_clustering(x.copy(), 0.9, 12, device="cpu")
    assert np.array_equal(a, b) and len(set(a)) == 3


def test_windows_follow_inference_slide():
    """10 s windows every 1 s; a trailing partial window is zero-padded (`Inference.slide`)."""
    pipe = P.SpeakerDiarization(None, None, device="cpu")
    w = pipe.windows(torch.arange(16000 * 12, dtype=torch.float32))          # 12 s -> windows at 0, 1, 2 s
    assert w.shape == (3, 160000) and float(w[2, 0]) == 32000.0
    w = pipe.windows(torch.ones(16000 * 12 + 8000))                         # + 0.5 s -> a padded 4th window
    assert w.shape == (4, 160000) and float(w[3, -1]) == 0.0 and float(w[3, 0]) == 1.0
    w = pipe.windows(torch.ones(5 * 16000))                                 # shorter than one window: one padded window
    assert w.shape == (1, 160000) and float(w[0, 5 * 16000 - 1]) == 1.0 and float(w[0, 5 * 16000]) == 0.0


def test_embedding_masks_exclude_overlap_unless_too_short():
    """`get_embeddings`: frames where two local speakers are active are dropped from a speaker's pooling mask, unless
    that leaves no more than `min_num_frames` clean frames — then the full mask is used."""
    seen = {}

    def emb(chunks, weights):
        seen["w"] = weights.clone()
        return torch.zeros(weights.shape[0], weights.shape[1], 4)

    pipe = P.SpeakerDiarization(None, emb, device="cpu")
    b = np.zeros((1, 589, 3), np.float32)
    b[0, 0:300, 0] = 1          # speaker 0: 300 frames, 100 of them overlapped by speaker 1
    b[0, 200:300, 1] = 1        # speaker 1: ONLY overlapped frames -> no clean frame -> falls back to its full mask
    b[0, 400:500, 2] = 1        # speaker 2: clean
    pipe.get_embeddings(torch.zeros(1, 160000), b)
    w = seen["w"][0].numpy()
    assert w.shape == (3, 589)
    assert w[0].sum() == 200 and w[0, 200:300].sum() == 0
    assert w[1].sum() == 100 and w[1, 200:300].sum() == 100
    assert w[2].sum() == 100
