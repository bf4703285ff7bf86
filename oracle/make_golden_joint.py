"""ORACLE tooling (test infrastructure): pins `joint_decoding` (time-synchronous joint CTC / attention beam search,
asr/wenet/transformer/search.py:450-496 + espnet/beam_search_timesync.py) against the LIVE reference.

The reference hard-codes sos=10000 (search.py:480), so the synthetic model needs V = 10001: the
small test shape with the full-size vocabulary.  Runs the reference's ASRModel.decode(['joint_decoding'], ...) on a
5.3 s synthetic recording and stores tokens / score / start times / token confidences in tests/golden/joint.json, and
checks on the spot that the oracle restatement (oracle/search_ref.joint_decoding) reproduces them.
Run from the repo root:  python oracle/make_golden_joint.py
"""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refimport  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SHAPE = dict(d=128, heads=2, ff=256, blocks=3, kernel=15, vocab=10001, dec_ff=256, dec_blocks=3, r_dec_blocks=3, emb_len=2)
CASES = [dict(name="causal_ln", causal=True, cnn_module_norm="layer_norm", model_seed=11, blank_rate=0.6),
         dict(name="sym_bn", causal=False, cnn_module_norm="batch_norm", model_seed=12, blank_rate=0.6)]
# a random decoder gives every token ~log(1/V) = -9.2, so the length bonus must outweigh (1 - ctc_weight) * 9.2 for
# non-empty hypotheses to survive (a trained decoder does that by itself)
SETTINGS = [dict(beam_size=4, ctc_weight=0.9, length_penalty=1.5), dict(beam_size=10, ctc_weight=0.5, length_penalty=5.0),
            dict(beam_size=10, ctc_weight=0.3, length_penalty=0.0)]


def main():
    sys.path.insert(0, ROOT)
    from reverb_b200 import synth
    from oracle import model_ref, pipeline_ref, search_ref
    wenet = refimport.import_reference()
    out = {"torch": torch.__version__, "shape": SHAPE, "audio_seconds": 5.3, "audio_seed": 99, "chunk_size": 260,
           "batch_size": 2, "verbatimicity": 1.0, "settings": SETTINGS, "cases": {}}
    for case in CASES:
        d = tempfile.mkdtemp()
        synth.write_model_dir(d, shape=SHAPE, causal=case["causal"], cnn_module_norm=case["cnn_module_norm"],
                              seed=case["model_seed"], blank_rate=case["blank_rate"])
        wav = synth.write_wav(os.path.join(d, "joint.wav"), synth.synth_audio(out["audio_seconds"], seed=out["audio_seed"]))
        m = wenet.load_model(d)
        orc = pipeline_ref.OracleASR(d)
        feats = m.compute_feats(wav, num_mel_bins=80, frame_length=25, frame_shift=10)
        cat = torch.tensor([1.0, 0.0])
        rec = {k: case[k] for k in ("causal", "cnn_module_norm", "model_seed", "blank_rate")}
        rec["runs"] = []
        with torch.no_grad():
            for st in SETTINGS:
                batches = []
                for fb, fl in m.feats_batcher(feats, out["chunk_size"], out["batch_size"]):
                    res = m.model.decode(["joint_decoding"], fb, fl, st["beam_size"], ctc_weight=st["ctc_weight"],
                                         length_penalty=st["length_penalty"], cat_embs=cat, blank_id=m.blank_id,
                                         infos={"tasks": ["transcribe"], "langs": ["en"]})["joint_decoding"]
                    # the oracle restatement on the oracle's own encoder / CTC tensors (bit-identical to the reference's)
                    enc, enc_lens, _ = orc.forward_encoder(fb, fl, cat)
                    ctc = model_ref.ctc_logprobs(enc, orc.sd)

                    def row(b, prefix):
                        mem = enc[b:b + 1, :int(enc_lens[b])]
                        return model_ref.decoder_step_logp(mem, enc_lens[b:b + 1], torch.tensor([prefix]), orc.sd, orc.cfg, cat)[0]
                    want = search_ref.joint_decoding(row, ctc, enc_lens, st["ctc_weight"], st["beam_size"], 1.5, st["length_penalty"])
                    for r, w in zip(res, want):
                        assert list(r.tokens) == list(w.tokens) and r.times == w.times, (r.tokens, w.tokens)
                        assert abs(r.score - w.score) < 1e-4 * max(1.0, abs(r.score)), (r.score, w.score)
                        assert max([abs(a - b) for a, b in zip(r.tokens_confidence, w.tokens_confidence)] + [0.0]) < 1e-5
                    batches.append([dict(tokens=list(map(int, r.tokens)), score=float(r.score), times=list(map(int, r.times)),
                                         tokens_confidence=[float(c) for c in r.tokens_confidence]) for r in res])
                rec["runs"].append(batches)
                print(case["name"], st, [[len(x["tokens"]) for x in b] for b in batches])
        out["cases"][case["name"]] = rec
    with open(os.path.join(GOLDEN, "joint.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
