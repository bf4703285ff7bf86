"""Hot-path parity on the GPU: CUDA path (through the C ABI) vs the committed reference fixtures
(tests/golden, produced by the live reference) and vs the CPU oracle on the same seeded inputs.

Tolerances (stated): the searches consume identical fp32 log-probs and must
be bit-exact in tokens / times (scores: 1e-9 relative, CUDA fp64 exp/log vs glibc).  The encoder runs
its GEMMs with bf16 operands and fp32 accumulation, so encoder_out / log-probs are compared with an
RMS-relative tolerance instead.
"""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def asr(model_dirs):
    import reverb_b200
    return {n: reverb_b200.load_model(d) for n, (d, _) in model_dirs.items()}


def _rel_rms(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-12))


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_searches_bit_exact_on_recorded_ctc_probs(asr, golden_cases, case):
    """greedy + prefix beam on the reference's own ctc_probs: tokens, n-best, times identical."""
    meta, arr = golden_cases[case]
    eng = asr[case].engine
    for bi, batch in enumerate(meta["batches"]):
        logp = torch.from_numpy(arr[f"ctc_probs_{bi}"]).cuda()
        lens = arr[f"enc_lens_{bi}"]
        val, idx = eng.logp_topk(logp, 10)
        # top-k itself: same values / indices as torch.topk on the recorded tensor
        tv, ti = logp.topk(10, dim=2)
        assert torch.equal(val, tv)
        assert torch.equal(idx.long(), ti)
        greedy = eng.greedy_search(idx, lens, 0)
        for b, g in enumerate(batch["ctc_greedy_search"]):
            assert greedy[b] == g["tokens"]
        pb = eng.prefix_beam_search(val, idx, lens, 10, 0)
        for b, g in enumerate(batch["ctc_prefix_beam_search"]):
            nbest, scores, times = pb[b]
            assert [list(h) for h in nbest] == g["nbest"]
            assert times == g["nbest_times"]
            np.testing.assert_allclose(scores, g["nbest_scores"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_fbank_and_encoder_vs_reference_fixture(asr, golden_cases, model_dirs, case):
    meta, arr = golden_cases[case]
    m = asr[case]
    feats = m.compute_feats(model_dirs[case][1], num_mel_bins=80, frame_length=25, frame_shift=10)
    np.testing.assert_allclose(feats[0].cpu().numpy(), arr["feats"], rtol=0, atol=2e-3)
    # encoder on the REFERENCE's features (identical fbank input), bf16-GEMM tolerance
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    ref_feats = torch.from_numpy(arr["feats"]).unsqueeze(0).cuda()
    for bi, (fb, fl) in enumerate(m.feats_batcher(ref_feats, meta["chunk_size"], meta["batch_size"])):
        enc, enc_lens = m.model._forward_encoder(fb, fl, cat)
        assert enc_lens.tolist() == arr[f"enc_lens_{bi}"].tolist()
        ref = arr[f"enc_out_{bi}"]
        got = enc.cpu().numpy()
        for b in range(ref.shape[0]):
            n = int(enc_lens[b])
            rr = _rel_rms(got[b, :n], ref[b, :n])
            print(f"[{case}] batch {bi} utt {b}: encoder_out rel-rms vs fp32 reference fixture = {rr:.2e}")
            assert rr < 6e-3
        logp = m.model.ctc_logprobs(enc).cpu().numpy()
        refp = arr[f"ctc_probs_{bi}"]
        for b in range(ref.shape[0]):
            n = int(enc_lens[b])
            # log-probs on the entries that matter (p > e^-12).  The synthetic CTC head is scaled x6
            # (logit sigma ~3.5, reverb_b200/synth.py), which amplifies the bf16 encoder error by the same factor:
            # stated tolerance 0.12 abs (max over ~7k entries), 0.025 RMS (measured 0.06 / 0.015).
            sel = refp[b, :n] > -12
            diff = logp[b, :n][sel] - refp[b, :n][sel]
            print(f"[{case}] batch {bi} utt {b}: log-prob max abs diff {np.abs(diff).max():.3f}, "
                  f"rms {np.sqrt((diff.astype(np.float64) ** 2).mean()):.4f}, "
                  f"argmax agreement {(logp[b, :n].argmax(-1) == refp[b, :n].argmax(-1)).mean():.3f}")
            assert np.abs(diff).max() < 0.12
            assert np.sqrt((diff.astype(np.float64) ** 2).mean()) < 0.025
            assert (logp[b, :n].argmax(-1) == refp[b, :n].argmax(-1)).mean() > 0.98


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_relpos_in_projection_epilogue_equals_the_separate_kernel(asr, golden_cases, case):
    """The rel-pos key transform (K'' = k + pos, key bias u.k + v.pos) runs in the [q; k; v] GEMM epilogue; RVB_RELPOS=prep
    keeps the separate relpos_prep kernel.  Same bf16 roundings, fp32 sums in a different order: encoder_out must agree
    far inside the bf16 tolerance."""
    meta, arr = golden_cases[case]
    m = asr[case]
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    ref_feats = torch.from_numpy(arr["feats"]).unsqueeze(0).cuda()
    fb, fl = next(iter(m.feats_batcher(ref_feats, meta["chunk_size"], meta["batch_size"])))
    enc_fused, _ = m.model._forward_encoder(fb, fl, cat)
    enc_fused = enc_fused.clone()
    os.environ["RVB_RELPOS"] = "prep"
    try:
        enc_prep, _ = m.model._forward_encoder(fb, fl, cat)
        enc_prep = enc_prep.clone()
    finally:
        del os.environ["RVB_RELPOS"]
    rr = _rel_rms(enc_fused.float().cpu().numpy(), enc_prep.float().cpu().numpy())
    print(f"[{case}] fused vs separate rel-pos: encoder_out rel-rms {rr:.2e}")
    assert rr < 5e-4


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_encoder_and_decoder_vs_bf16_emulating_oracle(asr, golden_cases, model_dirs, case):
    """Tight check of the kernels' logic: against the oracle with EMULATE_BF16 (it rounds to bf16 exactly where the
    engine stores bf16, everything else fp32) the only difference left is accumulation order, so the tolerance is
    ~5x tighter than against the pure-fp32 reference: encoder rel-RMS < 2.5e-3, log-probs 0.1 abs, decoder 0.03.
    (This test is what exposed the causal left-pad semantics of the conv module: pad frames are GLU(bias), not 0.)"""
    from oracle import model_ref, pipeline_ref, search_ref
    meta, arr = golden_cases[case]
    m = asr[case]
    orc = pipeline_ref.OracleASR(model_dirs[case][0])
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    ref_feats = torch.from_numpy(arr["feats"]).unsqueeze(0)
    model_ref.EMULATE_BF16 = True
    try:
        for bi, (fb, fl) in enumerate(orc.feats_batcher(ref_feats, meta["chunk_size"], meta["batch_size"])):
            with torch.no_grad():
                want, want_lens, _ = orc.forward_encoder(fb, fl, cat)
                want_logp = model_ref.ctc_logprobs(want, orc.sd)
            enc, enc_lens = m.model._forward_encoder(fb.cuda(), fl, cat)
            logp = m.model.ctc_logprobs(enc).cpu().numpy()
            got = enc.cpu().numpy()
            worst = 0.0
            for b in range(fb.shape[0]):
                n = int(enc_lens[b])
                worst = max(worst, _rel_rms(got[b, :n], want[b, :n].numpy()))
                sel = want_logp[b, :n].numpy() > -12
                assert np.abs(logp[b, :n][sel] - want_logp[b, :n].numpy()[sel]).max() < 0.1
            print(f"[{case}] encoder rel-rms vs bf16-emulating oracle: {worst:.2e}; "
                  f"vs fp32 reference: {_rel_rms(got[0, :int(enc_lens[0])], arr[f'enc_out_{bi}'][0, :int(enc_lens[0])]):.2e}")
            assert worst < 2.5e-3
            # decoder on the fp32 reference encoder_out / n-best
            g = meta["batches"][bi]["ctc_prefix_beam_search"]
            nbest = [[tuple(h) for h in r["nbest"]] for r in g]
            encr = torch.from_numpy(arr[f"enc_out_{bi}"])
            lens = arr[f"enc_lens_{bi}"]
            l2r, _ = m.engine.rescoring_scores(encr.cuda(), lens, nbest, cat, 0.0)
            for b, hyps in enumerate(nbest):
                ys, ylens = search_ref.rescoring_inputs(hyps, orc.sos, orc.eos)
                mem = encr[b, :int(lens[b])].unsqueeze(0).repeat(len(hyps), 1, 1)
                with torch.no_grad():
                    dec = torch.log_softmax(model_ref.decoder_forward(mem, ys, ylens, orc.sd, orc.cfg, "left_decoder", cat), -1)
                for i, h in enumerate(hyps):
                    U = len(h)
                    want_s = [float(dec[i, j, h[j]]) for j in range(U)] + [float(dec[i, U, orc.eos])]
                    np.testing.assert_allclose(l2r[b, i, :U + 1], want_s, rtol=0, atol=0.03)
    finally:
        model_ref.EMULATE_BF16 = False


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_rescoring_decoder_vs_oracle(asr, golden_cases, model_dirs, case):
    """teacher-forced decoder log-probs on the reference's encoder_out / n-best: vs the CPU oracle."""
    from oracle import model_ref, pipeline_ref, search_ref
    meta, arr = golden_cases[case]
    m = asr[case]
    orc = pipeline_ref.OracleASR(model_dirs[case][0])
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    rw = 0.3
    for bi, batch in enumerate(meta["batches"]):
        enc = torch.from_numpy(arr[f"enc_out_{bi}"])
        lens = arr[f"enc_lens_{bi}"]
        nbest = [[tuple(h) for h in g["nbest"]] for g in batch["ctc_prefix_beam_search"]]
        l2r, r2l = m.engine.rescoring_scores(enc.cuda(), lens, nbest, cat, rw)
        for b, hyps in enumerate(nbest):
            ys, ylens = search_ref.rescoring_inputs(hyps, orc.sos, orc.eos)
            mem = enc[b, :int(lens[b])].unsqueeze(0).repeat(len(hyps), 1, 1)
            dec = torch.log_softmax(model_ref.decoder_forward(mem, ys, ylens, orc.sd, orc.cfg, "left_decoder", cat), -1)
            rys = model_ref.reverse_hyps(ys, ylens, orc.eos)
            rdec = torch.log_softmax(model_ref.decoder_forward(mem, rys, ylens, orc.sd, orc.cfg, "right_decoder", cat), -1)
            for i, h in enumerate(hyps):
                U = len(h)
                want = [float(dec[i, j, h[j]]) for j in range(U)] + [float(dec[i, U, orc.eos])]
                rwant = [float(rdec[i, U - 1 - j, h[j]]) for j in range(U)] + [float(rdec[i, U, orc.eos])]
                np.testing.assert_allclose(l2r[b, i, :U + 1], want, rtol=0, atol=0.15)
                np.testing.assert_allclose(r2l[b, i, :U + 1], rwant, rtol=0, atol=0.15)


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_decode_is_consistent_with_oracle_searches_on_gpu_logprobs(asr, golden_cases, model_dirs, case):
    """decode() end to end: the GPU searches / rescoring must equal the oracle's searches run on the SAME
    (GPU-produced) log-probs and decoder scores — i.e. the only deviation from the reference anywhere in the
    chain is the stated bf16 tolerance of encoder / decoder activations."""
    from oracle import search_ref
    from reverb_b200.search import rescoring_pick
    meta, arr = golden_cases[case]
    m = asr[case]
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    rw, cw = meta["reverse_weight"], meta["ctc_weight"]
    ref_feats = torch.from_numpy(arr["feats"]).unsqueeze(0).cuda()
    modes = ["ctc_greedy_search", "ctc_prefix_beam_search", "attention_rescoring"]
    for fb, fl in m.feats_batcher(ref_feats, meta["chunk_size"], meta["batch_size"]):
        # the flat rescoring decoder (one row per hypothesis and position) is the same function as
        # engine.rescoring_scores below; the default prefix-tree decoder agrees with it to rounding
        # (tests/test_gpu_parity.py::test_prefix_tree_rescoring_equals_the_flat_decoder)
        os.environ["RVB_RESCORE"] = "flat"
        try:
            got = m.model.decode(modes, fb, fl, 10, ctc_weight=cw, reverse_weight=rw, cat_embs=cat, blank_id=0)
        finally:
            del os.environ["RVB_RESCORE"]
        enc, enc_lens = m.model._forward_encoder(fb, fl, cat)
        logp = m.model.ctc_logprobs(enc).cpu()
        lens_t = torch.from_numpy(enc_lens.astype(np.int64))
        want_g = search_ref.ctc_greedy_search(logp, lens_t, 0)
        want_p = search_ref.ctc_prefix_beam_search(logp, lens_t, 10, 0)
        l2r, r2l = m.engine.rescoring_scores(enc, enc_lens, [w.nbest for w in want_p], cat, rw)
        for b in range(fb.shape[0]):
            assert got["ctc_greedy_search"][b].tokens == want_g[b].tokens
            assert got["ctc_greedy_search"][b].times is None                      # like the reference
            gp = got["ctc_prefix_beam_search"][b]
            assert [tuple(h) for h in gp.nbest] == [tuple(h) for h in want_p[b].nbest]
            assert gp.nbest_times == want_p[b].nbest_times
            np.testing.assert_allclose(gp.nbest_scores, want_p[b].nbest_scores, rtol=1e-9, atol=1e-9)
            assert gp.tokens == gp.nbest[0] and gp.times == gp.nbest_times[0] and gp.tokens_confidence is None
            want_r = rescoring_pick(want_p[b].nbest, want_p[b].nbest_scores, want_p[b].nbest_times, l2r[b],
                                    None if r2l is None else r2l[b], cw, rw)
            gr = got["attention_rescoring"][b]
            assert tuple(gr.tokens) == tuple(want_r.tokens) and gr.times == want_r.times
            assert abs(gr.score - want_r.score) < 1e-4 and abs(gr.confidence - want_r.confidence) < 1e-6
            np.testing.assert_allclose(gr.tokens_confidence, want_r.tokens_confidence, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_attention_mode_decoder_steps_and_search(asr, golden_cases, model_dirs, case):
    """`attention` decode mode (search.py:251-360).  (1) the GPU decoder step (left decoder, last position,
    log_softmax + top-k) agrees with the oracle's forward_one_step restatement on the same running hypotheses;
    (2) ASRModel.decode(['attention']) equals the oracle's beam search driven by the GPU step function (the host
    bookkeeping is the only other ingredient); (3) tokens vs the live-reference fixture: equal whenever the bf16
    step log-probs do not reorder a near-tie — reported, with a floor on the agreement."""
    import json as _json
    from oracle import model_ref, pipeline_ref, search_ref
    meta, arr = golden_cases[case]
    m = asr[case]
    gold = _json.load(open(os.path.join(os.path.dirname(__file__), "golden", "attention_mode.json")))["cases"][case]
    orc = pipeline_ref.OracleASR(model_dirs[case][0])
    sd, cfg = orc.sd, orc.cfg
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    ref_feats = torch.from_numpy(arr["feats"]).unsqueeze(0).cuda()
    N, agree, total = 10, 0, 0
    for bi, (fb, fl) in enumerate(m.feats_batcher(ref_feats, meta["chunk_size"], meta["batch_size"])):
        enc, enc_lens = m.model._forward_encoder(fb, fl, cat)
        B, Tp, d = enc.shape
        # (1) step numerics on a few hand-made prefixes
        rng = np.random.default_rng(bi)
        for L in (1, 2, 5):
            hyps = rng.integers(1, 100, size=(B * N, L)).astype(np.int64)
            hyps[:, 0] = m.model.sos
            val, idx = m.engine.decoder_step_topk(enc, enc_lens, hyps, N, cat, N)
            mem = enc.cpu().unsqueeze(1).repeat(1, N, 1, 1).view(B * N, Tp, d)
            mem_lens = torch.from_numpy(enc_lens.astype(np.int64)).view(-1, 1).repeat(1, N).view(-1)
            want = model_ref.decoder_step_logp(mem, mem_lens, torch.from_numpy(hyps), sd, cfg, cat)
            got_at = torch.gather(want, 1, torch.from_numpy(idx.astype(np.int64)))
            assert float((got_at - torch.from_numpy(val)).abs().max()) < 0.15
            assert float((want.topk(N).values - torch.from_numpy(val)).abs().max()) < 0.15
        # (2) full search: GPU decode == oracle bookkeeping over the GPU step function
        for lp in (0.0, 0.6):
            os.environ["RVB_ATTENTION_STEP"] = "recompute"      # the cache-free step, same function as `step` below
            try:
                got = m.model.decode(["attention"], fb, fl, N, length_penalty=lp, cat_embs=cat, blank_id=0)["attention"]
            finally:
                del os.environ["RVB_ATTENTION_STEP"]

            def step(hyps):
                v, i = m.engine.decoder_step_topk(enc, enc_lens, hyps.numpy(), N, cat, N)
                return torch.from_numpy(v), torch.from_numpy(i.astype(np.int64))
            want = search_ref.attention_beam_search(step, B, Tp, N, m.model.sos, m.model.eos, lp)
            assert [list(r.tokens) for r in got] == [list(r.tokens) for r in want]
            assert all(r.times is None for r in got)
            for b in range(B):
                total += 1
                agree += int(list(got[b].tokens) == gold[f"length_penalty_{lp}"][bi][b])
    assert agree >= 0.5 * total, (agree, total)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_kv_cached_decoder_step_equals_prefix_recompute(model_dirs, golden_cases, precision):
    """rvb_decoder_cache_* (one new position per step, per-layer key / value cache, beam reordering by gather) against
    rvb_decoder_step_topk (recomputes the prefix): the same log_softmax top-k, step after step, under random beam
    re-rankings.  Then `attention` mode end to end: in the fp32-accurate mode the tokens of the live reference
    (tests/golden/attention_mode.json) exactly."""
    import json as _json
    import reverb_b200
    case = "causal_ln"
    meta, arr = golden_cases[case]
    m = reverb_b200.load_model(model_dirs[case][0], precision=precision)
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    feats = torch.from_numpy(arr["feats"]).unsqueeze(0).cuda()
    fb, fl = next(iter(m.feats_batcher(feats, meta["chunk_size"], meta["batch_size"])))
    enc, enc_lens = m.model._forward_encoder(fb, fl, cat)
    B, N = enc.shape[0], 4
    S = B * N
    rng = np.random.default_rng(5)
    hyps = np.full((S, 1), m.model.sos, dtype=np.int64)
    m.engine.decoder_cache_begin(enc, enc_lens, N, 12, cat)
    tol = 2e-3 if precision == "fp32" else 0.12
    try:
        parents = None
        for step in range(8):
            v_c, i_c = m.engine.decoder_cache_step(hyps[:, -1], parents, 6)
            v_r, i_r = m.engine.decoder_step_topk(enc, enc_lens, hyps, N, cat, 6)
            assert float(np.abs(v_c - v_r).max()) < tol, (step, float(np.abs(v_c - v_r).max()))
            if precision == "fp32":
                assert (i_c == i_r).mean() > 0.97
            # re-rank: every utterance's hypotheses pick random parents among its own N, and extend them
            parents = (np.arange(S) // N) * N + rng.integers(0, N, size=S)
            new_tok = i_r[parents, rng.integers(0, 6, size=S)]
            hyps = np.concatenate([hyps[parents], new_tok[:, None]], axis=1)
    finally:
        m.engine.decoder_cache_end()
    gold = _json.load(open(os.path.join(os.path.dirname(__file__), "golden", "attention_mode.json")))["cases"][case]
    agree = total = 0
    for bi, (fb, fl) in enumerate(m.feats_batcher(feats, meta["chunk_size"], meta["batch_size"])):
        for lp in (0.0, 0.6):
            got = m.model.decode(["attention"], fb, fl, 10, length_penalty=lp, cat_embs=cat, blank_id=0)["attention"]
            for b, r in enumerate(got):
                total += 1
                agree += int(list(r.tokens) == gold[f"length_penalty_{lp}"][bi][b])
    print(f"[attention mode, KV-cached, {precision}] hypotheses identical to the live reference: {agree}/{total}")
    assert agree == total or precision == "bf16"


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_transcribe_api_surface(asr, golden_cases, model_dirs, case):
    """Public API: CTM / TXT strings, chunk offsets, error behaviour of the reference."""
    meta, arr = golden_cases[case]
    m = asr[case]
    wav = model_dirs[case][1]
    kw = dict(verbatimicity=meta["verbatimicity"], chunk_size=meta["chunk_size"], batch_size=meta["batch_size"],
              reverse_weight=meta["reverse_weight"])
    out = m.transcribe_modes(wav, ["ctc_prefix_beam_search", "attention_rescoring"], format="ctm", **kw)
    n_ref = len(meta["transcribe"]["attention_rescoring.ctm"].split("\n"))
    for text in out:
        lines = text.split("\n")
        assert 0.5 * n_ref < len(lines) < 2 * n_ref
        prev_start = -1.0
        for ln in lines:
            f = ln.split(" ")
            assert len(f) == 6 and f[0] == "golden.wav" and f[1] == "0" and float(f[3]) >= 0
            assert float(f[2]) >= prev_start                                        # chunk offsets monotone
            prev_start = float(f[2])
    assert all(ln.endswith(" 0.00") for ln in out[0].split("\n"))                   # prefix beam: no confidences
    got_words = [ln.split(" ")[4] for ln in out[1].split("\n")]
    txt = m.transcribe(wav, mode="attention_rescoring", format="txt", **kw)
    assert txt.split(" ") == got_words
    # batch size must not change the result (chunks are independent)
    kw1 = dict(kw, batch_size=1)
    assert m.transcribe(wav, mode="attention_rescoring", format="ctm", **kw1) == out[1]
    with pytest.raises(ValueError):
        m.transcribe(wav, format="json")
    with pytest.raises((AssertionError, TypeError)):
        m.transcribe(wav, mode="ctc_greedy_search")          # reference quirk 1: greedy has no times
    with pytest.raises(IndexError):
        m.transcribe(wav, mode="joint_decoding")              # reference quirk 3: sos=10000 is hard-coded, V = 101 here


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_bounded_context_encoder_vs_reference_fixture(asr, golden_cases, case):
    """decoding_chunk_size > 0: chunk-masked attention (utils/mask.py:88-197) inside the tcgen05 attention kernel;
    encoder_out vs the live-reference fixture with the same bf16 tolerance as the full-context encoder, and the
    chunked output must differ from the full-context one (the mask is really applied)."""
    import json as _json
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    gold = _json.load(open(os.path.join(gdir, "chunked.json")))
    arr_c = dict(np.load(os.path.join(gdir, "chunked.npz")))
    meta, arr = golden_cases[case]
    m = asr[case]
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    ref_feats = torch.from_numpy(arr["feats"]).unsqueeze(0).cuda()
    for cs, left in gold["settings"]:
        for bi, (fb, fl) in enumerate(m.feats_batcher(ref_feats, meta["chunk_size"], meta["batch_size"])):
            enc, lens = m.model._forward_encoder(fb, fl, cat, decoding_chunk_size=cs, num_decoding_left_chunks=left)
            full, _ = m.model._forward_encoder(fb, fl, cat)
            want = arr_c[f"{case}_c{cs}_l{left}_enc_{bi}"]
            for b in range(fb.shape[0]):
                n = int(lens[b])
                assert _rel_rms(enc[b, :n].cpu().numpy(), want[b, :n]) < 6e-3
                assert _rel_rms(full[b, :n].cpu().numpy(), want[b, :n]) > 2e-2
            res = m.model.decode(["ctc_greedy_search"], fb, fl, 10, decoding_chunk_size=cs,
                                 num_decoding_left_chunks=left, cat_embs=cat, blank_id=0)
            assert len(res["ctc_greedy_search"]) == fb.shape[0]


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_simulate_streaming_equals_the_cache_based_reference(asr, golden_cases, case):
    """simulate_streaming: the reference's CACHE-based chunk-by-chunk encoder (encoder.forward_chunk_by_chunk,
    encoder.py:341-402; golden from the live reference, oracle/make_golden_streaming.py) vs this engine's single masked
    pass — attention cache == chunk mask, causal cnn cache == left context, non-causal conv == chunk-local conv."""
    import json as _json
    import reverb_b200
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    gold = _json.load(open(os.path.join(gdir, "streaming.json")))
    arr_s = dict(np.load(os.path.join(gdir, "streaming.npz")))
    meta, arr = golden_cases[case]
    m = asr[case]
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    feats = torch.from_numpy(arr["feats"][:gold["frames"]]).unsqueeze(0).cuda()
    lens = torch.tensor([gold["frames"]], dtype=torch.int32)
    for cs, left in gold["settings"]:
        want = arr_s[f"{case}_c{cs}_l{left}"]
        enc, enc_lens = m.model._forward_encoder(feats, lens, cat, cs, left, simulate_streaming=True)
        assert enc.shape[1] == want.shape[0] == int(enc_lens[0])
        rr = _rel_rms(enc[0].cpu().numpy(), want)
        print(f"[{case}] simulate_streaming chunk {cs} left {left}: rel-rms vs the cache-based reference {rr:.2e}")
        assert rr < 6e-3
        res = m.model.decode(["ctc_greedy_search", "attention_rescoring"], feats, lens, 10, decoding_chunk_size=cs,
                             num_decoding_left_chunks=left, simulate_streaming=True, cat_embs=cat, ctc_weight=0.1, blank_id=0)
        assert len(res["attention_rescoring"]) == 1


def test_compute_feats_resamples_non_16k_audio_on_the_gpu(asr, model_dirs, tmp_path):
    """cli/reverb.py:120-138: a WAV at another rate is resampled to 16 kHz (torchaudio Resample semantics) before
    fbank; here both steps run on the GPU and must agree with the oracle chain resample_ref -> fbank_np."""
    from oracle import fbank_np, resample_ref
    from reverb_b200 import synth
    m = asr["causal_ln"]
    for rate in (8000, 44100):
        pcm = synth.synth_audio(1.3, seed=31)[: int(1.3 * rate)]
        wav = synth.write_wav(str(tmp_path / f"r{rate}.wav"), pcm, sample_rate=rate)
        feats = m.compute_feats(wav, num_mel_bins=80, frame_length=25, frame_shift=10)
        want_wave = resample_ref.resample(torch.from_numpy(pcm.astype(np.float32)).unsqueeze(0), rate, 16000)[0].numpy()
        want = fbank_np.fbank(want_wave)
        assert tuple(feats.shape) == (1,) + want.shape
        np.testing.assert_allclose(feats[0].cpu().numpy(), want, rtol=0, atol=5e-3)


def test_launch_counter_and_no_cpu_path():
    from reverb_b200.engine import launch_count
    assert launch_count() > 0


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_decode_with_context_graph(asr, golden_cases, case):
    """`ASRModel.decode(context_graph=...)`: the biased prefix search (host, bit-exact vs the live reference on recorded
    log-probs: tests/test_context_biasing.py) runs on the GPU's top-k and feeds the rescoring decoder."""
    from reverb_b200.context_graph import ContextGraph
    from reverb_b200.search import ctc_prefix_beam_search_biased
    meta, arr = golden_cases[case]
    m = asr[case]
    cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
    ref_feats = torch.from_numpy(arr["feats"]).unsqueeze(0).cuda()
    fb, fl = next(iter(m.feats_batcher(ref_feats, meta["chunk_size"], meta["batch_size"])))
    beam = meta["beam_size"]
    modes = ["ctc_prefix_beam_search", "attention_rescoring"]
    plain = m.model.decode(modes, fb, fl, beam, ctc_weight=0.5, cat_embs=cat)
    best = list(plain["ctc_prefix_beam_search"][0].tokens)
    assert len(best) >= 4
    phrase = best[1:4]
    graph = ContextGraph(token_lists=[phrase], context_score=3.0)
    biased = m.model.decode(modes, fb, fl, beam, ctc_weight=0.5, cat_embs=cat, context_graph=graph)
    # (after `finalize` the reported scores carry no matched bonus — search.py:228-233 replaces the context score by minus
    #  the bonus of an unfinished match — so nothing is asserted about score ordering against the unbiased run)
    # equals the host search run directly on the GPU's top-k
    enc, enc_lens = m.model._forward_encoder(fb, fl, cat)
    val, idx, _ = m.model.engine.ctc_topk(enc, beam)
    want = ctc_prefix_beam_search_biased(val.cpu().numpy(), idx.cpu().numpy(), enc_lens, beam, graph, 0)
    for r, w in zip(biased["ctc_prefix_beam_search"], want):
        assert [list(h) for h in r.nbest] == [list(h) for h in w.nbest] and r.nbest_scores == w.nbest_scores
    # rescoring consumed the biased n-best
    for r, pr in zip(biased["attention_rescoring"], biased["ctc_prefix_beam_search"]):
        assert tuple(r.tokens) in [tuple(h) for h in pr.nbest]
