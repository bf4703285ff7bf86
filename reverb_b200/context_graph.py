"""Context biasing graph for CTC prefix beam search (host side), compiled to flat arrays.

Behaviour = `asr/wenet/utils/context_graph.py` of the reference (ContextGraph :104-265, tokenize :24-58): an
Aho-Corasick automaton over the token sequences of the biasing phrases.  Every node carries the bonus accumulated from
the root (`depth * context_score`), a fail link and the summed bonus of the phrases that END at it or at a node on its
output chain; the search feeds it one token at a time (`forward_one_step`) and closes a hypothesis with `finalize`.
The reference's observable quirks are kept: when a token does not extend the current node, the fail walk stops at the
root WITHOUT retrying the root's own children unless the loop lands there (:217-225, same in the construction :168-176);
a phrase that ends on an already existing node does not mark it as a phrase end (:150-161).

Representation: states are plain integers (0 = root) indexing parallel lists — `children[s]` (token -> state),
`fail[s]`, `bonus[s]`, `emit[s]` — instead of linked node objects; `ASRModel.decode(context_graph=...)` and
`reverb_b200.search.ctc_prefix_beam_search_biased` only use `root`, `forward_one_step` and `finalize`, so the
reference's own `ContextGraph` object can be passed as well.  The reverb CLI never builds a graph (cli/reverb.py:227).
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

_CJK = re.compile(r"([一-鿿])")


def tokenize(context_list_path: str, symbol_table: Dict[str, int], bpe_model: Optional[str] = None) -> List[List[int]]:
    """One biasing phrase per line -> token ids.  With a sentencepiece model: upper-cased text, CJK characters on their
    own, everything else through `encode_as_pieces` (text/tokenize_utils.py:19-60); without: one symbol per character,
    space written as the word-boundary mark.  Symbols missing from the table become <unk> if the table has it and are
    dropped otherwise."""
    encode = None
    if bpe_model is not None:
        import sentencepiece as spm
        sp = spm.SentencePieceProcessor()
        sp.load(bpe_model)

        def encode(text: str) -> List[str]:
            pieces: List[str] = []
            for part in _CJK.split(text.upper()):
                if part.strip():
                    pieces.extend([part] if _CJK.fullmatch(part) else sp.encode_as_pieces(part))
            return pieces
    unk = symbol_table.get("<unk>")
    phrases: List[List[int]] = []
    with open(context_list_path, "r") as f:
        for line in f:
            text = line.strip()
            symbols = encode(text) if encode else [("▁" if ch == " " else ch) for ch in text]
            ids = [symbol_table.get(sym, unk) for sym in symbols]
            phrases.append([i for i in ids if i is not None])
    return phrases


class ContextGraph:
    root = 0

    def __init__(self, context_list_path: Optional[str] = None, symbol_table: Optional[Dict[str, int]] = None,
                 bpe_model: Optional[str] = None, context_score: float = 6.0,
                 token_lists: Optional[Iterable[Sequence[int]]] = None):
        """Positional arguments as in the reference (`context_list_path, symbol_table, bpe_model, context_score`);
        `token_lists` builds the graph from token ids directly."""
        self.context_score = context_score
        self.context_list = ([list(t) for t in token_lists] if token_lists is not None
                             else tokenize(context_list_path, symbol_table or {}, bpe_model))
        self.children: List[Dict[int, int]] = [{}]
        self.token: List[int] = [-1]
        self.bonus: List[float] = [0.0]       # accumulated bonus root -> node ("node_score")
        self.emit: List[float] = [0.0]        # bonus of the phrases recognised on arrival ("output_score")
        self.ends: List[bool] = [False]
        self.fail: List[int] = [0]
        self._insert_phrases()
        self._link()

    @property
    def num_nodes(self) -> int:
        return len(self.token) - 1

    def _insert_phrases(self) -> None:
        for phrase in self.context_list:
            s = 0
            for pos, tok in enumerate(phrase):
                nxt = self.children[s].get(tok)
                if nxt is None:
                    nxt = len(self.token)
                    last = pos == len(phrase) - 1
                    depth_bonus = self.bonus[s] + self.context_score
                    self.children[s][tok] = nxt
                    self.children.append({})
                    self.token.append(tok)
                    self.bonus.append(depth_bonus)
                    self.emit.append(depth_bonus if last else 0)
                    self.ends.append(last)
                    self.fail.append(0)
                s = nxt

    def _fallback(self, start: int, tok: int) -> int:
        """Walk fail links from `start` until a node with a `tok` child is found or the root is reached; take that child
        if there is one.  (The root's children are only consulted when the walk ends on the root.)"""
        s = start
        while tok not in self.children[s]:
            s = self.fail[s]
            if s == 0:
                break
        return self.children[s].get(tok, s)

    def _link(self) -> None:
        order = list(self.children[0].values())          # breadth first; depth-1 nodes fail to the root
        head = 0
        while head < len(order):
            parent = order[head]
            head += 1
            for tok, node in self.children[parent].items():
                f = self.fail[parent]
                self.fail[node] = self.children[f][tok] if tok in self.children[f] else self._fallback(self.fail[f], tok)
                # nearest phrase end on the fail chain contributes its (already complete) output bonus
                out = self.fail[node]
                while not self.ends[out]:
                    out = self.fail[out]
                    if out == 0:
                        out = -1
                        break
                if out >= 0:
                    self.emit[node] += self.emit[out]
                order.append(node)

    # -- queries ------------------------------------------------------------------------------------------------------
    def forward_one_step(self, state: int, token: int) -> Tuple[float, int]:
        nxt = self.children[state].get(token)
        if nxt is not None:
            gained = self.context_score
        else:
            nxt = self._fallback(self.fail[state], token)
            gained = self.bonus[nxt] - self.bonus[state]
        return gained + self.emit[nxt], nxt

    def finalize(self, state: int) -> Tuple[float, int]:
        """Take back the bonus of a match that did not complete; the next state is the root."""
        return -self.bonus[state], 0
