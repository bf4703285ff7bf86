"""ORACLE tooling: golden vectors for bounded-context decoding (decoding_chunk_size > 0, utils/mask.py:88-197) from
the LIVE reference: the synthetic models of tests/golden/{causal_ln,sym_bn}.json, encoder_out + greedy / prefix-beam
tokens for two (chunk, left) settings -> tests/golden/chunked.npz + chunked.json.
Run from the repo root:  python oracle/make_golden_chunked.py"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refimport  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SETTINGS = [(16, -1), (8, 2)]


def main():
    sys.path.insert(0, ROOT)
    from reverb_b200 import synth
    wenet = refimport.import_reference()
    arrays, meta_out = {}, {"settings": SETTINGS, "cases": {}}
    for name in ("causal_ln", "sym_bn"):
        meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
        d = tempfile.mkdtemp()
        synth.write_model_dir(d, causal=meta["causal"], cnn_module_norm=meta["cnn_module_norm"],
                              seed=meta["model_seed"], blank_rate=meta["blank_rate"])
        wav = synth.write_wav(os.path.join(d, "golden.wav"), synth.synth_audio(meta["audio_seconds"], seed=meta["audio_seed"]))
        m = wenet.load_model(d)
        feats = m.compute_feats(wav, num_mel_bins=80, frame_length=25, frame_shift=10)
        cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
        case = {}
        with torch.no_grad():
            for cs, left in SETTINGS:
                batches = []
                for bi, (fb, fl) in enumerate(m.feats_batcher(feats, meta["chunk_size"], meta["batch_size"])):
                    enc, mask = m.model._forward_encoder(fb, fl, decoding_chunk_size=cs, num_decoding_left_chunks=left,
                                                         cat_embs=cat)
                    res = m.model.decode(["ctc_greedy_search", "ctc_prefix_beam_search"], fb, fl, 10,
                                         decoding_chunk_size=cs, num_decoding_left_chunks=left, cat_embs=cat,
                                         blank_id=m.blank_id, infos={"tasks": ["transcribe"], "langs": ["en"]})
                    arrays[f"{name}_c{cs}_l{left}_enc_{bi}"] = enc.numpy()
                    batches.append({"greedy": [list(map(int, r.tokens)) for r in res["ctc_greedy_search"]],
                                    "prefix": [list(map(int, r.tokens)) for r in res["ctc_prefix_beam_search"]]})
                case[f"c{cs}_l{left}"] = batches
        meta_out["cases"][name] = case
        print(name, "done")
    np.savez_compressed(os.path.join(GOLDEN, "chunked.npz"), **arrays)
    json.dump(meta_out, open(os.path.join(GOLDEN, "chunked.json"), "w"), indent=1)
    print({k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
