"""ORACLE tooling (test infrastructure): pins context biasing — `ContextGraph` (asr/wenet/utils/context_graph.py) and the
`context_graph is not None` branches of `ctc_prefix_beam_search` (asr/wenet/transformer/search.py:124-248) — against the
LIVE reference.

Input = the CTC log-probabilities the reference itself recorded for the two small fixtures (tests/golden/{causal_ln,
sym_bn}.npz, ctc_probs_0).  The biasing phrases are token-id lists (the reference's `build_graph` takes them directly;
its text front-end needs a sentencepiece model that the synthetic model directory does not have): two 3-token spans cut
from the reference's own unbiased 2nd / 4th best hypotheses — so that biasing visibly reorders the n-best — plus an
overlapping pair that exercises the fail / output arcs.  Stored in tests/golden/context.json: the phrases, the biased
search results (n-best tokens, scores, times) and, for random token streams, the (score, node id) trace of
`forward_one_step` / `finalize`.
Run from the repo root:  python oracle/make_golden_context.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refimport  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def ref_graph(cg_mod, token_lists, score):
    g = cg_mod.ContextGraph.__new__(cg_mod.ContextGraph)
    g.context_score = score
    g.context_list = [list(t) for t in token_lists]
    g.num_nodes = 0
    g.root = cg_mod.ContextState(id=0, token=-1, token_score=0, node_score=0, output_score=0, is_end=False)
    g.root.fail = g.root
    g.build_graph(g.context_list)
    return g


def main():
    refimport.import_reference()
    from wenet.transformer import search as rsearch
    from wenet.utils import context_graph as cg_mod
    out = {"torch": torch.__version__, "cases": {}}
    rng = np.random.default_rng(5)
    for name in ("causal_ln", "sym_bn"):
        arr = np.load(os.path.join(GOLDEN, name + ".npz"))
        meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
        probs = torch.from_numpy(arr["ctc_probs_0"])
        lens = torch.from_numpy(arr["enc_lens_0"])
        beam = int(meta["beam_size"])
        plain = rsearch.ctc_prefix_beam_search(probs, lens, beam, None, 0)
        phrases = []
        for r, which in ((plain[0], 1), (plain[-1], min(3, beam - 1))):
            hyp = list(r.nbest[which])
            if len(hyp) >= 4:
                s = len(hyp) // 2
                phrases.append(hyp[s - 1:s + 2])
        a, b2, c = (int(x) for x in rng.integers(1, probs.shape[2] - 1, 3))
        phrases += [[a, b2, c], [b2, c], [c]]                      # overlapping suffixes: fail + output arcs
        runs = []
        for score in (3.0, 6.0):
            g = ref_graph(cg_mod, phrases, score)
            res = rsearch.ctc_prefix_beam_search(probs, lens, beam, g, 0)
            runs.append({"context_score": score,
                         "results": [{"nbest": [list(map(int, h)) for h in r.nbest],
                                      "nbest_scores": [float(s) for s in r.nbest_scores],
                                      "nbest_times": [list(map(int, t)) for t in r.nbest_times]} for r in res],
                         "changed_vs_plain": [list(map(int, r.nbest[0])) != list(map(int, p.nbest[0])) or
                                              [list(map(int, h)) for h in r.nbest] != [list(map(int, h)) for h in p.nbest]
                                              for r, p in zip(res, plain)]})
        # automaton traces
        g = ref_graph(cg_mod, phrases, 6.0)
        vocab_hot = sorted({t for ph in phrases for t in ph})
        traces = []
        for _ in range(6):
            stream = [int(rng.choice(vocab_hot)) if rng.random() < 0.7 else int(rng.integers(1, probs.shape[2]))
                      for _ in range(25)]
            st = g.root
            steps = []
            for tok in stream:
                sc, st = g.forward_one_step(st, tok)
                steps.append([float(sc), int(st.id)])
            fin, _ = g.finalize(st)
            traces.append({"stream": stream, "steps": steps, "finalize": float(fin)})
        out["cases"][name] = {"beam_size": beam, "phrases": phrases, "runs": runs, "num_nodes": int(g.num_nodes),
                              "traces": traces}
        print(name, "phrases", phrases, "changed:", [r["changed_vs_plain"] for r in runs])
    with open(os.path.join(GOLDEN, "context.json"), "w") as f:
        json.dump(out, f)
    print("wrote tests/golden/context.json")


if __name__ == "__main__":
    main()
