"""ORACLE tooling (test infrastructure): pins the cache-based streaming simulation
(BaseEncoder.forward_chunk_by_chunk, asr/wenet/transformer/encoder.py:341-402) against the LIVE reference.

For the two synthetic models of tests/golden/{causal_ln,sym_bn}.json it runs the reference's
`model.encoder.forward_chunk_by_chunk(feats[:, :T], chunk, left, cat_embs)` (the encoder method: ASRModel.decode drops
cat_embs on this path and asserts for LSL models, asr_model.py:299-303) and the oracle restatement
(model_ref.encoder_forward_chunk_by_chunk), checks they agree, and stores encoder_out in tests/golden/streaming.npz.
Run from the repo root:  python oracle/make_golden_streaming.py
"""
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
SETTINGS = [(16, -1), (8, 2), (5, 0)]     # (decoding_chunk_size, num_decoding_left_chunks)
FRAMES = 397                              # not a multiple of the stride: exercises the short last window


def main():
    sys.path.insert(0, ROOT)
    from reverb_b200 import synth
    from oracle import model_ref, pipeline_ref
    wenet = refimport.import_reference()
    arrays, meta_out = {}, {"torch": torch.__version__, "settings": SETTINGS, "frames": FRAMES}
    for name in ("causal_ln", "sym_bn"):
        meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
        d = tempfile.mkdtemp()
        synth.write_model_dir(d, causal=meta["causal"], cnn_module_norm=meta["cnn_module_norm"],
                              seed=meta["model_seed"], blank_rate=meta["blank_rate"])
        m = wenet.load_model(d)
        orc = pipeline_ref.OracleASR(d)
        feats = torch.from_numpy(np.load(os.path.join(GOLDEN, name + ".npz"))["feats"])[:FRAMES].unsqueeze(0)
        cat = torch.tensor([meta["verbatimicity"], 1.0 - meta["verbatimicity"]])
        with torch.no_grad():
            for cs, left in SETTINGS:
                ref, _ = m.model.encoder.forward_chunk_by_chunk(feats, cs, left, cat_embs=cat)
                mine = model_ref.encoder_forward_chunk_by_chunk(feats, orc.sd, orc.cfg, cat, cs, left)
                err = float((ref - mine).abs().max())
                print(name, cs, left, tuple(ref.shape), "oracle vs live reference max abs diff", err)
                assert ref.shape == mine.shape and err < 2e-5
                arrays[f"{name}_c{cs}_l{left}"] = ref[0].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "streaming.npz"), **arrays)
    with open(os.path.join(GOLDEN, "streaming.json"), "w") as f:
        json.dump(meta_out, f, indent=1)


if __name__ == "__main__":
    main()
