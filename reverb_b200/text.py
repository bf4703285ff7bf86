"""Token id <-> piece mapping (the only tokenizer functionality the inference path needs).

Mirrors the parts of asr/wenet/text/{char_tokenizer,rev_bpe_tokenizer}.py and
asr/wenet/utils/file_utils.py:61-68 that `ReverbASR` touches: the symbol table
(`<piece> <id>` per line) and `detokenize(ids) -> (text, pieces)`.  Text -> tokens needs
sentencepiece and is only used for training; it is loaded lazily exactly like the
reference does (rev_bpe_tokenizer.py:35-39).
"""
from __future__ import annotations

from typing import Dict, List, Tuple


def read_symbol_table(path: str) -> Dict[str, int]:
    table: Dict[str, int] = {}
    with open(path, "r", encoding="utf8") as f:
        for line in f:
            parts = line.strip().split()
            if len(parts) != 2:
                raise ValueError(f"bad symbol table line in {path!r}: {line!r}")
            table[parts[0]] = int(parts[1])
    return table


class PieceTokenizer:
    def __init__(self, symbol_table_path: str, bpe_model_path: str = None, unk: str = "<unk>",
                 connect_symbol: str = ""):
        self.symbol_table = read_symbol_table(symbol_table_path)
        self.id2piece = {i: p for p, i in self.symbol_table.items()}
        self.unk = unk
        self.connect_symbol = connect_symbol
        self._bpe_path = bpe_model_path
        self._sp = None

    def vocab_size(self) -> int:
        return len(self.symbol_table)

    def ids2tokens(self, ids: List[int]) -> List[str]:
        return [self.id2piece[int(i)] for i in ids]

    def tokens2text(self, tokens: List[str]) -> str:
        return self.connect_symbol.join(tokens).replace("▁", " ").strip()

    def detokenize(self, ids: List[int]) -> Tuple[str, List[str]]:
        pieces = self.ids2tokens(ids)
        return self.tokens2text(pieces), pieces

    def text2tokens(self, line: str) -> List[str]:
        if self._sp is None:
            import sentencepiece as spm
            self._sp = spm.SentencePieceProcessor()
            self._sp.load(self._bpe_path)
        line = line.strip().replace("<sw>", "").replace("  ", " ").strip().replace("<unk>", "<unknown>")
        return self._sp.encode(line, out_type=str)

    def tokens2ids(self, tokens: List[str]) -> List[int]:
        unk_id = self.symbol_table.get(self.unk)
        return [self.symbol_table.get(t, unk_id) for t in tokens]


def init_tokenizer(configs: Dict) -> PieceTokenizer:
    tc = configs["tokenizer_conf"]
    kind = configs.get("tokenizer", "char")
    if kind not in ("rev_bpe", "bpe", "char"):
        raise NotImplementedError(f"tokenizer {kind!r} is not supported by reverb_b200")
    return PieceTokenizer(tc["symbol_table_path"], tc.get("bpe_path"), connect_symbol=tc.get("connect_symbol", ""))


def get_blank_id(configs: Dict, symbol_table: Dict[str, int]) -> int:
    """asr/wenet/utils/ctc_utils.py:164-178."""
    cc = configs.setdefault("ctc_conf", {})
    if "<blank>" in symbol_table:
        if "ctc_blank_id" in cc:
            assert cc["ctc_blank_id"] == symbol_table["<blank>"]
        else:
            cc["ctc_blank_id"] = symbol_table["<blank>"]
    else:
        assert "ctc_blank_id" in cc, "PLZ set ctc_blank_id in yaml"
    return cc["ctc_blank_id"]
