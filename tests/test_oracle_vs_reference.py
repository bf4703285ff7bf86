"""Re-pins the oracle and the synthetic-model generator against the LIVE reference whenever
/root/reference is mounted (authoring container).  Skipped on the GPU box, where the committed
fixtures (tests/golden, produced by oracle/make_golden.py from the same reference) take over."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import refimport

pytestmark = pytest.mark.skipif(not refimport.available(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def wenet_ref():
    warnings.filterwarnings("ignore")
    return refimport.import_reference()


def test_synthetic_state_dict_is_strictly_loadable(wenet_ref, model_dirs):
    d, _ = model_dirs["causal_ln"]
    m = wenet_ref.load_model(d)
    ref_sd = m.model.state_dict()
    sd = torch.load(os.path.join(d, "synth.pt"))
    assert set(ref_sd.keys()) == set(sd.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref_sd[k].shape), k
    assert type(m.model.encoder.encoders[0]).__name__ == "LanguageSpecificConformerEncoderLayer"
    assert type(m.model.decoder).__name__ == "LanguageSpecificBiTransformerDecoder"


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_oracle_equals_live_reference(wenet_ref, model_dirs, golden_cases, case):
    from oracle import pipeline_ref
    meta, _ = golden_cases[case]
    d, wav = model_dirs[case]
    m = wenet_ref.load_model(d)
    orc = pipeline_ref.OracleASR(d)
    f_ref = m.compute_feats(wav, num_mel_bins=80, frame_length=25, frame_shift=10)
    assert (f_ref - orc.compute_feats(wav)).abs().max().item() < 5e-4
    cat = torch.tensor([0.25, 0.75])
    modes = ["ctc_greedy_search", "ctc_prefix_beam_search", "attention_rescoring"]
    for fb, fl in m.feats_batcher(f_ref, 350, 2):
        with torch.no_grad():
            want = m.model.decode(modes, fb, fl, 7, ctc_weight=0.3, reverse_weight=0.5, cat_embs=cat,
                                  infos={"tasks": ["transcribe"], "langs": ["en"]})
            enc_ref, _ = m.model._forward_encoder(fb, fl, cat_embs=cat)
        got = orc.decode(modes, fb, fl, 7, ctc_weight=0.3, reverse_weight=0.5, cat_embs=cat, return_intermediates=True)
        assert torch.equal(enc_ref, got["_encoder_out"])
        for b in range(fb.shape[0]):
            assert want["ctc_greedy_search"][b].tokens == got["ctc_greedy_search"][b].tokens
            a, c = want["ctc_prefix_beam_search"][b], got["ctc_prefix_beam_search"][b]
            assert a.nbest == c.nbest and a.nbest_scores == c.nbest_scores and a.nbest_times == c.nbest_times
            a, c = want["attention_rescoring"][b], got["attention_rescoring"][b]
            assert tuple(a.tokens) == tuple(c.tokens) and float(a.score) == float(c.score)
            assert a.confidence == c.confidence and a.tokens_confidence == c.tokens_confidence


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_oracle_attention_mode_and_bounded_context_equal_live_reference(wenet_ref, model_dirs, golden_cases, case):
    """The later restatements — `attention` decode mode (search.py:251-360) and decoding_chunk_size > 0
    (utils/mask.py:88-197) — against the live reference with settings the committed fixtures do not use."""
    from oracle import pipeline_ref
    d, wav = model_dirs[case]
    m = wenet_ref.load_model(d)
    orc = pipeline_ref.OracleASR(d)
    feats = m.compute_feats(wav, num_mel_bins=80, frame_length=25, frame_shift=10)
    cat = torch.tensor([0.4, 0.6])
    for fb, fl in m.feats_batcher(feats, 300, 2):
        with torch.no_grad():
            want = m.model.decode(["attention"], fb, fl, 5, length_penalty=0.3, cat_embs=cat,
                                  infos={"tasks": ["transcribe"], "langs": ["en"]})
            enc_ref, _ = m.model._forward_encoder(fb, fl, decoding_chunk_size=12, num_decoding_left_chunks=1, cat_embs=cat)
            want_c = m.model.decode(["ctc_prefix_beam_search"], fb, fl, 6, decoding_chunk_size=12, num_decoding_left_chunks=1,
                                    cat_embs=cat, infos={"tasks": ["transcribe"], "langs": ["en"]})
        got = orc.decode(["attention"], fb, fl, 5, cat_embs=cat, length_penalty=0.3)
        assert [list(r.tokens) for r in got["attention"]] == [list(r.tokens) for r in want["attention"]]
        got_c = orc.decode(["ctc_prefix_beam_search"], fb, fl, 6, cat_embs=cat, return_intermediates=True,
                           decoding_chunk_size=12, num_decoding_left_chunks=1)
        assert torch.equal(enc_ref, got_c["_encoder_out"])
        for a, c in zip(want_c["ctc_prefix_beam_search"], got_c["ctc_prefix_beam_search"]):
            assert a.nbest == c.nbest and a.nbest_scores == c.nbest_scores and a.nbest_times == c.nbest_times


def test_oracle_resample_equals_torchaudio():
    import torchaudio
    from oracle import resample_ref
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 7777, generator=g) * 3000
    for rate in (8000, 32000, 44100):
        want = torchaudio.transforms.Resample(orig_freq=rate, new_freq=16000)(x)
        assert torch.equal(resample_ref.resample(x, rate, 16000), want)


def test_host_post_processing_equals_live_reference(wenet_ref, golden_cases, model_dirs):
    """reverb_b200's ctc_align / CTM rendering vs the reference's, on the reference's own hypotheses."""
    from wenet.bin.ctc_align import adjust_model_time_offset as ref_adjust, ctc_align as ref_align
    from reverb_b200 import ctc_align as mine
    from reverb_b200.text import PieceTokenizer
    meta, _ = golden_cases["causal_ln"]
    d, _ = model_dirs["causal_ln"]
    m = wenet_ref.load_model(d)
    tok = PieceTokenizer(os.path.join(d, "tk.units.txt"))
    for batch in meta["batches"]:
        for r in batch["attention_rescoring"]:
            a = ref_adjust(ref_align(r["tokens"], r["times"], r["tokens_confidence"], m.tokenizer, 40, 1230), 230)
            b = mine.adjust_model_time_offset(mine.ctc_align(r["tokens"], r["times"], r["tokens_confidence"], tok, 40, 1230), 230)
            assert a == b
