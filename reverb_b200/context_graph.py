"""Context biasing graph for CTC prefix beam search (host side).

Restates `asr/wenet/utils/context_graph.py` of the reference (ContextGraph :104-265, tokenize :24-58): an Aho-Corasick
automaton over the token sequences of the biasing phrases — a trie whose nodes carry the accumulated bonus
(`node_score` = depth * context_score), a fail arc and an output arc — queried one token at a time by the search
(`forward_one_step`) and closed with `finalize`.  Scores and tie-breaking follow the reference exactly, including its
quirks: the fail-arc walk stops at the root without retrying from it (:168-176, :217-225), and a fail transition earns
`node.node_score - state.node_score` plus the output score of the node reached.

Used by `reverb_b200.search.ctc_prefix_beam_search_biased` through `ASRModel.decode(context_graph=...)` — the
`context_graph` argument of the reference's `ASRModel.decode` (asr_model.py:331-350).  The reverb CLI itself never
builds one (cli/reverb.py:227 passes `context_graph=None`).
"""
from __future__ import annotations

import re
from collections import deque
from typing import Dict, Iterable, List, Optional, Sequence, Tuple


_CJK = re.compile(r"([\u4e00-\u9fff])")


def tokenize(context_list_path: str, symbol_table: Dict[str, int], bpe_model: Optional[str] = None) -> List[List[int]]:
    """One biasing phrase per line -> token ids: sentencepiece pieces when a BPE model is given, otherwise characters
    (space -> '▁'); unknown symbols map to <unk> when the table has one, else they are dropped
    (context_graph.py:24-58)."""
    sp = None
    if bpe_model is not None:
        import sentencepiece as spm
        sp = spm.SentencePieceProcessor()
        sp.load(bpe_model)
    out: List[List[int]] = []
    with open(context_list_path, "r") as f:
        for line in f:
            text = line.strip()
            if sp is not None:
                # text/tokenize_utils.py:19-20, 28-60: upper-case, CJK characters stand alone, the rest goes through the
                # sentencepiece model
                pieces = []
                for part in _CJK.split(text.upper()):
                    if not part.strip():
                        continue
                    if _CJK.fullmatch(part):
                        pieces.append(part)
                    else:
                        pieces.extend(sp.encode_as_pieces(part))
            else:
                pieces = ["▁" if ch == " " else ch for ch in text]
            ids = []
            for piece in pieces:
                if piece in symbol_table:
                    ids.append(symbol_table[piece])
                elif "<unk>" in symbol_table:
                    ids.append(symbol_table["<unk>"])
            out.append(ids)
    return out


class ContextState:
    """A trie node.  `token` of the root is -1."""
    __slots__ = ("id", "token", "token_score", "node_score", "output_score", "is_end", "next", "fail", "output")

    def __init__(self, id: int, token: int, token_score: float, node_score: float, output_score: float, is_end: bool):
        self.id = id
        self.token = token
        self.token_score = token_score
        self.node_score = node_score
        self.output_score = output_score
        self.is_end = is_end
        self.next: Dict[int, "ContextState"] = {}
        self.fail: Optional["ContextState"] = None
        self.output: Optional["ContextState"] = None


class ContextGraph:
    def __init__(self, context_list_path: Optional[str] = None, symbol_table: Optional[Dict[str, int]] = None,
                 bpe_model: Optional[str] = None, context_score: float = 6.0,
                 token_lists: Optional[Iterable[Sequence[int]]] = None):
        """Same positional arguments as the reference (`context_list_path, symbol_table, bpe_model, context_score`);
        `token_lists` builds the graph from token ids directly."""
        self.context_score = context_score
        if token_lists is not None:
            self.context_list = [list(t) for t in token_lists]
        else:
            self.context_list = tokenize(context_list_path, symbol_table or {}, bpe_model)
        self.num_nodes = 0
        self.root = ContextState(0, -1, 0, 0, 0, False)
        self.root.fail = self.root
        self._build(self.context_list)

    def _build(self, token_lists: List[List[int]]) -> None:
        for tokens in token_lists:
            node = self.root
            for i, tok in enumerate(tokens):
                nxt = node.next.get(tok)
                if nxt is None:
                    self.num_nodes += 1
                    end = i == len(tokens) - 1
                    score = node.node_score + self.context_score
                    nxt = ContextState(self.num_nodes, tok, self.context_score, score, score if end else 0, end)
                    node.next[tok] = nxt
                node = nxt           # NB (reference :150-161): a phrase that ends on an existing inner node does not mark it
        # fail / output arcs, breadth first
        queue = deque()
        for node in self.root.next.values():
            node.fail = self.root
            queue.append(node)
        while queue:
            cur = queue.popleft()
            for tok, node in cur.next.items():
                fail = cur.fail
                if tok in fail.next:
                    fail = fail.next[tok]
                else:
                    fail = fail.fail
                    while tok not in fail.next:
                        fail = fail.fail
                        if fail.token == -1:
                            break
                    if tok in fail.next:
                        fail = fail.next[tok]
                node.fail = fail
                out = node.fail
                while not out.is_end:
                    out = out.fail
                    if out.token == -1:
                        out = None
                        break
                node.output = out
                node.output_score += 0 if out is None else out.output_score
                queue.append(node)

    def forward_one_step(self, state: ContextState, token: int) -> Tuple[float, ContextState]:
        if token in state.next:
            node = state.next[token]
            score = node.token_score
        else:
            node = state.fail
            while token not in node.next:
                node = node.fail
                if node.token == -1:
                    break
            if token in node.next:
                node = node.next[token]
            score = node.node_score - state.node_score
        return score + node.output_score, node

    def finalize(self, state: ContextState) -> Tuple[float, ContextState]:
        return -state.node_score, self.root
