"""Host-side logic of the product (no GPU): chunking, word/CTM assembly, score combination, loaders."""
import json
import os

import numpy as np
import pytest
import torch


class _Cfg:
    pass


def _fake_asr():
    """ReverbASR without the engine: only the pure-host methods are exercised."""
    from reverb_b200.reverb import ReverbASR
    obj = ReverbASR.__new__(ReverbASR)
    obj.test_conf = {"fbank_conf": {"num_mel_bins": 80, "frame_length": 25, "frame_shift": 10}}
    return obj


@pytest.mark.parametrize("m,chunk,batch", [(1128, 400, 2), (898, 330, 3), (2051, 2051, 1), (5, 400, 2), (800, 400, 2)])
def test_feats_batcher_matches_oracle_restatement(m, chunk, batch):
    from oracle.pipeline_ref import OracleASR
    asr = _fake_asr()
    feats = torch.randn(1, m, 80)
    mine = list(asr.feats_batcher(feats, chunk, batch))
    ref = list(OracleASR.feats_batcher(feats, chunk, batch))
    assert len(mine) == len(ref)
    for (a, al), (b, bl) in zip(mine, ref):
        assert torch.equal(a, b) and al.tolist() == bl.tolist() and al.dtype == torch.int32


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_ctm_and_txt_rendering_equal_reference_strings(golden_cases, model_dirs, case):
    """get_output on the reference's hypotheses reproduces the reference's transcribe() strings exactly."""
    from reverb_b200.reverb import get_output
    from reverb_b200.search import DecodeResult
    from reverb_b200.text import PieceTokenizer
    meta, _ = golden_cases[case]
    tok = PieceTokenizer(os.path.join(model_dirs[case][0], "tk.units.txt"))
    for mode in ("ctc_prefix_beam_search", "attention_rescoring"):
        hyps = []
        for batch in meta["batches"]:
            for r in batch[mode]:
                hyps.append(DecodeResult(r["tokens"], r["score"], r["confidence"], r["tokens_confidence"], r["times"]))
        for fmt in ("ctm", "txt"):
            got = get_output(fmt, tok, "golden.wav", hyps, 230, meta["chunk_size"], 10, 40)
            assert got == meta["transcribe"][f"{mode}.{fmt}"]
    with pytest.raises(ValueError):
        get_output("json", tok, "x", [], 230, 400, 10, 40)


def test_ctc_align_requires_times_like_the_reference(model_dirs):
    from reverb_b200.ctc_align import adjust_model_time_offset, ctc_align
    from reverb_b200.text import PieceTokenizer
    tok = PieceTokenizer(os.path.join(model_dirs["causal_ln"][0], "tk.units.txt"))
    with pytest.raises((TypeError, AssertionError)):
        ctc_align([3, 4], None, None, tok, 40, 0)
    assert ctc_align([], [], None, tok, 40, 0) == []
    assert adjust_model_time_offset([], 0) is None          # reference quirk 4


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_rescoring_pick_float_semantics(golden_cases, model_dirs, case):
    """rescoring_pick on oracle decoder log-probs reproduces the reference's score / confidence bit for bit."""
    from oracle import model_ref, pipeline_ref, search_ref
    from reverb_b200.search import rescoring_pick
    meta, arr = golden_cases[case]
    orc = pipeline_ref.OracleASR(model_dirs[case][0])
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    rw, cw = meta["reverse_weight"], meta["ctc_weight"]
    for bi, batch in enumerate(meta["batches"]):
        enc = torch.from_numpy(arr[f"enc_out_{bi}"])
        lens = arr[f"enc_lens_{bi}"]
        for b, g in enumerate(batch["ctc_prefix_beam_search"]):
            hyps = [tuple(h) for h in g["nbest"]]
            ys, ylens = search_ref.rescoring_inputs(hyps, orc.sos, orc.eos)
            mem = enc[b, :int(lens[b])].unsqueeze(0).repeat(len(hyps), 1, 1)
            with torch.no_grad():
                dec = torch.log_softmax(model_ref.decoder_forward(mem, ys, ylens, orc.sd, orc.cfg, "left_decoder", cat), -1)
                rdec = None
                if rw > 0:
                    rys = model_ref.reverse_hyps(ys, ylens, orc.eos)
                    rdec = torch.log_softmax(model_ref.decoder_forward(mem, rys, ylens, orc.sd, orc.cfg, "right_decoder", cat), -1)
            L = ys.shape[1]
            l2r = np.zeros((len(hyps), L), np.float32)
            r2l = np.zeros((len(hyps), L), np.float32) if rdec is not None else None
            for i, h in enumerate(hyps):
                U = len(h)
                for j in range(U):
                    l2r[i, j] = dec[i, j, h[j]]
                    if r2l is not None:
                        r2l[i, j] = rdec[i, U - 1 - j, h[j]]
                l2r[i, U] = dec[i, U, orc.eos]
                if r2l is not None:
                    r2l[i, U] = rdec[i, U, orc.eos]
            got = rescoring_pick(hyps, g["nbest_scores"], g["nbest_times"], l2r, r2l, cw, rw)
            want = batch["attention_rescoring"][b]
            assert list(got.tokens) == want["tokens"] and got.times == want["times"]
            assert got.score == want["score"] and got.confidence == want["confidence"]
            assert got.tokens_confidence == want["tokens_confidence"]


def test_cmvn_loaders(tmp_path):
    from reverb_b200.cmvn import load_cmvn
    mean = np.array([1.0, -2.0, 0.5])
    var = np.array([4.0, 0.25, 1e-30])
    n = 50.0
    js = tmp_path / "cmvn.json"
    js.write_text(json.dumps({"mean_stat": (mean * n).tolist(), "var_stat": ((var + mean ** 2) * n).tolist(), "frame_num": n}))
    m, istd = load_cmvn(str(js), True)
    np.testing.assert_allclose(m, mean)
    np.testing.assert_allclose(istd[:2], 1 / np.sqrt(var[:2]))
    assert istd[2] == pytest.approx(1e10)                                  # variance floor 1e-20
    kd = tmp_path / "cmvn.kaldi"
    kd.write_text("[\n " + " ".join(str(x) for x in mean * n) + f" {n}\n " + " ".join(str(x) for x in (var + mean ** 2) * n) + " 0 ]\n")
    m2, istd2 = load_cmvn(str(kd), False)
    np.testing.assert_allclose(m2, m)
    np.testing.assert_allclose(istd2, istd)


def test_load_model_error_behaviour(tmp_path):
    import reverb_b200
    assert reverb_b200.get_available_models() == ["reverb_asr_v1"]
    with pytest.raises(ValueError):
        reverb_b200.load_model("no_such_model_name")
    import wenet
    assert wenet.load_model is reverb_b200.load_model
    if not torch.cuda.is_available():
        from reverb_b200 import synth
        d = synth.write_model_dir(str(tmp_path / "m"))
        with pytest.raises(RuntimeError, match="CUDA"):                      # no CPU fallback
            reverb_b200.load_model(d)


def test_cli_argument_surface():
    from reverb_b200.recognize_wav import get_args, main
    a = get_args(["--audio_file", "a.wav", "--result_dir", "out", "--model", "m"])
    assert a.modes == ["attention_rescoring"] and a.beam_size == 10 and a.chunk_size == 2051 and a.batch_size == 1
    assert a.ctc_weight == 0.1 and a.reverse_weight == 0.0 and a.verbatimicity == 1.0 and a.timings_adjustment == 230
    with pytest.raises(RuntimeError):
        main(["--audio_file", "a.wav", "--result_dir", "out"])              # neither --model nor config+checkpoint


def test_encoder_length_formula_matches_mask_slicing():
    from reverb_b200 import _lib
    lib = _lib.load()
    for T in (7, 8, 9, 10, 11, 330, 400, 2051, 2998):
        Tp = ((T - 1) // 2 - 1) // 2
        assert lib.rvb_encoder_out_frames(T) == Tp
        for ln in (0, 1, 6, 7, 8, 10, 11, 12, T - 1, T):
            if ln > T:
                continue
            mask = (torch.arange(T) < ln)[None, None, :]
            want = int(mask[:, :, 2::2][:, :, 2::2].sum())
            assert lib.rvb_encoder_out_len(ln, T) == want, (T, ln)


def test_rescoring_pick_batch_equals_per_utterance_pick():
    """The array path used by decode() (fused native search + rescoring) picks exactly like `rescoring_pick`."""
    from reverb_b200.search import rescoring_pick, rescoring_pick_batch
    rng = np.random.default_rng(0)
    B, N, L = 5, 10, 12
    olen = np.zeros((B, N, 2), np.int32)
    toks = rng.integers(1, 90, (B, N, L)).astype(np.int32)
    tims = np.sort(rng.integers(0, 99, (B, N, L)), axis=2).astype(np.int32)
    nhyp = np.array([10, 10, 7, 1, 10], np.int32)
    l2r = np.zeros((B, N, L + 1), np.float32)
    r2l = np.zeros((B, N, L + 1), np.float32)
    for b in range(B):
        for i in range(N):
            U = int(rng.integers(0, L + 1))
            olen[b, i] = (U, U)
            l2r[b, i, :U + 1] = -rng.random(U + 1).astype(np.float32) * 3
            r2l[b, i, :U + 1] = -rng.random(U + 1).astype(np.float32) * 3
    sc = -rng.random((B, N)) * 20
    for rw in (0.0, 0.3):
        got = rescoring_pick_batch(toks, tims, olen, sc, nhyp, l2r, r2l if rw > 0 else None, 0.1, rw)
        for b in range(B):
            n = int(nhyp[b])
            hyps = [tuple(toks[b, i, :olen[b, i, 0]].tolist()) for i in range(n)]
            want = rescoring_pick(hyps, list(sc[b, :n]), [tims[b, i, :olen[b, i, 1]].tolist() for i in range(n)],
                                  l2r[b], r2l[b] if rw > 0 else None, 0.1, rw)
            g = got[b]
            assert tuple(g.tokens) == tuple(want.tokens) and g.score == want.score and g.times == want.times
            assert g.confidence == want.confidence
            np.testing.assert_allclose(g.tokens_confidence, want.tokens_confidence, rtol=1e-15)


def test_attention_beam_search_host_logic_equals_oracle_restatement():
    """`attention` mode bookkeeping (numpy, reverb_b200/search.py) vs the oracle's torch restatement of
    search.py:251-360 on the same synthetic step function (deterministic pseudo-decoder with an <eos> that becomes
    likely after a few tokens), several beam sizes / length penalties / batch sizes."""
    from oracle import search_ref
    from reverb_b200.search import attention_beam_search
    V, sos, eos = 50, 49, 49

    def make_step(seed, N):
        def logp_of(hyps):
            rows = []
            for h in np.asarray(hyps):
                key = int(np.sum(np.asarray(h, dtype=np.int64) * np.arange(1, len(h) + 1))) * 7919 + seed + 131 * len(h)
                g = np.random.default_rng(key % (2 ** 32))
                x = g.standard_normal(V).astype(np.float32) * 2.0
                x[eos] += 0.9 * (len(h) - 3)
                rows.append(torch.log_softmax(torch.from_numpy(x), dim=0))
            return torch.stack(rows)
        prev = {}

        def n_step(hyps, parents=None):
            # the product passes `parents` (row of the previous call's hyps each row extends): check the contract
            hyps = np.asarray(hyps)
            if parents is not None:
                assert np.array_equal(prev["hyps"][np.asarray(parents)], hyps[:, :-1])
            prev["hyps"] = hyps.copy()
            return tuple(t.numpy() for t in logp_of(hyps).topk(N))
        return (lambda hyps: logp_of(hyps).topk(N)), n_step
    for seed, B, N, lp, maxlen in [(0, 3, 10, 0.0, 12), (1, 2, 4, 0.6, 9), (2, 1, 1, 0.0, 6), (3, 4, 10, 1.0, 5)]:
        t_step, n_step = make_step(seed, N)
        want = search_ref.attention_beam_search(t_step, B, maxlen, N, sos, eos, lp)
        got = attention_beam_search(n_step, B, maxlen, N, sos, eos, lp)
        assert [list(r.tokens) for r in got] == [list(r.tokens) for r in want]


def test_attention_context_follows_add_optional_chunk_mask_rules():
    """Which (chunk, left) the encoder applies (utils/mask.py:126-197): dynamic-chunk configs use the call's
    decoding_chunk_size, static-chunk configs their own size, everything else full context."""
    from reverb_b200.asr_model import ASRModel
    m = ASRModel.__new__(ASRModel)
    m.configs = {"encoder_conf": {"use_dynamic_chunk": True}}
    assert m.attention_context(-1, -1) == (-1, -1)
    assert m.attention_context(16, -1) == (16, -1)
    assert m.attention_context(8, 2) == (8, 2)
    with pytest.raises(AssertionError):
        m.attention_context(0, -1)
    m.configs = {"encoder_conf": {"use_dynamic_chunk": False, "static_chunk_size": 12}}
    assert m.attention_context(-1, 3) == (12, 3)
    assert m.attention_context(5, -1) == (12, -1)
    m.configs = {"encoder_conf": {}}
    assert m.attention_context(16, 2) == (-1, -1)


def test_resample_table_shapes_and_lengths():
    from reverb_b200.resample import resampled_length, sinc_resample_kernel
    for rate, (o, n) in {8000: (1, 2), 48000: (3, 1), 44100: (441, 160), 22050: (441, 320)}.items():
        kern, orig, new, width = sinc_resample_kernel(rate, 16000)
        assert (orig, new) == (o, n) and kern.shape == (new, 2 * width + orig) and kern.dtype == np.float32
        assert resampled_length(1000, orig, new) == -(-1000 * new // orig)
