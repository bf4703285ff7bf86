"""Context biasing: `reverb_b200.context_graph.ContextGraph` and the host prefix beam search with a
context graph, against tests/golden/context.json — produced by the LIVE reference's `ContextGraph` / `ctc_prefix_beam_search`
on its own recorded CTC log-probabilities (oracle/make_golden_context.py).  No GPU: the search consumes per-frame top-k
arrays, here taken with torch.topk from the recorded log-probabilities exactly like the reference does."""
import json
import os

import numpy as np
import pytest
import torch

from reverb_b200.context_graph import ContextGraph, tokenize
from reverb_b200.search import ctc_prefix_beam_search_biased

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "context.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_context_graph_automaton_traces_equal_the_reference(golden, case):
    g = golden["cases"][case]
    graph = ContextGraph(token_lists=g["phrases"], context_score=6.0)
    assert graph.num_nodes == g["num_nodes"]
    for tr in g["traces"]:
        st = graph.root
        for tok, (score, node_id) in zip(tr["stream"], tr["steps"]):
            sc, st = graph.forward_one_step(st, tok)
            assert sc == score and st == node_id
        assert graph.finalize(st)[0] == tr["finalize"]


@pytest.mark.parametrize("case", ["causal_ln", "sym_bn"])
def test_biased_prefix_beam_search_equals_the_reference(golden, case):
    g = golden["cases"][case]
    arr = np.load(os.path.join(GOLDEN, case + ".npz"))
    probs = torch.from_numpy(arr["ctc_probs_0"])
    lens = arr["enc_lens_0"]
    beam = g["beam_size"]
    val, idx = probs.topk(beam, dim=2)
    for run in g["runs"]:
        graph = ContextGraph(token_lists=g["phrases"], context_score=run["context_score"])
        res = ctc_prefix_beam_search_biased(val.numpy(), idx.numpy(), lens, beam, graph, 0)
        assert any(run["changed_vs_plain"]), "the fixture must actually be affected by the biasing"
        for r, want in zip(res, run["results"]):
            assert [list(h) for h in r.nbest] == want["nbest"]
            assert [list(t) for t in r.nbest_times] == want["nbest_times"]
            np.testing.assert_allclose(r.nbest_scores, want["nbest_scores"], rtol=1e-12, atol=0)
            assert list(r.tokens) == want["nbest"][0] and r.score == r.nbest_scores[0]


def test_tokenize_characters_and_unknowns(tmp_path):
    table = {"<blank>": 0, "<unk>": 1, "a": 2, "b": 3, "▁": 4}
    p = tmp_path / "ctx.txt"
    p.write_text("ab a\nzb\n")
    assert tokenize(str(p), table) == [[2, 3, 4, 2], [1, 3]]
    del table["<unk>"]
    assert tokenize(str(p), table) == [[2, 3, 4, 2], [3]]
    g = ContextGraph(str(p), table, None, 2.0)
    sc, st = g.forward_one_step(g.root, 2)
    assert sc == 2.0 and g.bonus[st] == 2.0
