"""`ASRModel.decode` — the operator boundary the kernels sit behind.

Same signature, argument meaning and result types as the reference's
`ASRModel.decode` (asr/wenet/transformer/asr_model.py:331-432); the body drives the native
plan (reverb_b200/engine.py) instead of a torch.nn graph:

    encoder (csrc/engine.cu) -> CTC head + top-k (csrc/ctc.cu) -> greedy / prefix beam (GPU)
    -> teacher-forced decoder over the n-best (GPU) -> score combination (host, search.py)
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from .engine import Engine
from .search import (DecodeResult, attention_beam_search, greedy_results, prefix_beam_results, rescoring_pick,
                     rescoring_pick_batch)

SUPPORTED_METHODS = ("attention", "ctc_greedy_search", "ctc_prefix_beam_search", "attention_rescoring")


class ASRModel:
    def __init__(self, engine: Engine, configs: Dict, vocab_size: int):
        self.engine = engine
        self.configs = configs
        self.vocab_size = vocab_size
        st = (configs.get("tokenizer_conf") or {}).get("special_tokens")
        # asr_model.py:79-82: eos is the same id as sos
        self.sos = vocab_size - 1 if st is None else st.get("<sos>", vocab_size - 1)
        self.eos = vocab_size - 1 if st is None else st.get("<eos>", vocab_size - 1)
        self.ignore_id = -1
        ds = configs.get("dataset_conf", {})
        self.lsl_enc = self.lsl_dec = bool(ds.get("pass_cat_emb", False))
        self.reverse_weight = configs.get("model_conf", {}).get("reverse_weight", 0.0)

    def sos_symbol(self) -> int:
        return self.sos

    def eos_symbol(self) -> int:
        return self.eos

    def eval(self):
        return self

    def to(self, device):
        return self

    # -- pieces of decode(), exposed for tests / profiling ---------------------------------------
    def attention_context(self, decoding_chunk_size: int, num_decoding_left_chunks: int):
        """The (chunk, left) pair add_optional_chunk_mask (utils/mask.py:126-197) would apply for this model's
        encoder_conf; (-1, -1) = full context."""
        ec = self.configs.get("encoder_conf", {})
        if ec.get("use_dynamic_chunk", False):
            if decoding_chunk_size < 0:
                return -1, -1
            if decoding_chunk_size > 0:
                return int(decoding_chunk_size), int(num_decoding_left_chunks)
            raise AssertionError("decoding_chunk_size == 0 selects the random training chunks")   # asr_model.py:377
        if ec.get("static_chunk_size", 0) > 0:
            return int(ec["static_chunk_size"]), int(num_decoding_left_chunks)
        return -1, -1

    def _forward_encoder(self, speech: torch.Tensor, speech_lengths, cat_embs=None, decoding_chunk_size: int = -1,
                         num_decoding_left_chunks: int = -1):
        """-> (encoder_out (B, T', d) fp32 cuda, encoder_lens np.int32 (B,))."""
        lens = speech_lengths.detach().cpu().numpy() if torch.is_tensor(speech_lengths) else np.asarray(speech_lengths)
        chunk, left = self.attention_context(decoding_chunk_size, num_decoding_left_chunks)
        return self.engine.forward_encoder(speech, lens, cat_embs, chunk, left)

    def ctc_logprobs(self, encoder_out: torch.Tensor, blank_penalty: float = 0.0, blank_id: int = 0) -> torch.Tensor:
        return self.engine.ctc_topk(encoder_out, 1, blank_penalty, blank_id, want_logp=True)[2]

    @torch.no_grad()
    def decode(self, methods: List[str], speech: torch.Tensor, speech_lengths: torch.Tensor, beam_size: int,
               decoding_chunk_size: int = -1, num_decoding_left_chunks: int = -1, ctc_weight: float = 0.0,
               simulate_streaming: bool = False, reverse_weight: float = 0.0, context_graph=None,
               blank_id: int = 0, blank_penalty: float = 0.0, length_penalty: float = 0.0,
               infos: Optional[Dict[str, List[str]]] = None, cat_embs: Optional[torch.Tensor] = None,
               cv=None, cv_lengths=None) -> Dict[str, List[DecodeResult]]:
        assert speech.shape[0] == speech_lengths.shape[0]
        assert decoding_chunk_size != 0
        if simulate_streaming and decoding_chunk_size > 0:
            # encoder.forward_chunk_by_chunk (encoder.py:231-402): chunk-by-chunk with att / cnn caches
            raise NotImplementedError("reverb_b200: simulate_streaming (cache-based chunk-by-chunk encoding) is not built; "
                                      "decoding_chunk_size > 0 without it applies the same bounded attention context")
        if context_graph is not None:
            raise NotImplementedError("reverb_b200: context biasing is out of scope (SURVEY.md §2)")
        unknown = [m for m in methods if m not in SUPPORTED_METHODS]
        if unknown:
            raise NotImplementedError(f"reverb_b200: decoding method(s) {unknown} are not built yet (SURVEY.md §8f)")
        if not speech.is_cuda:
            speech = speech.to(self.engine.device, non_blocking=True)
        speech = speech.to(torch.float32)
        encoder_out, encoder_lens = self._forward_encoder(speech, speech_lengths, cat_embs, decoding_chunk_size,
                                                          num_decoding_left_chunks)
        need_beam = "ctc_prefix_beam_search" in methods or "attention_rescoring" in methods
        k = beam_size if need_beam else 1
        topk_val, topk_idx, _ = self.engine.ctc_topk(encoder_out, k, blank_penalty, blank_id)
        results: Dict[str, List[DecodeResult]] = {}
        if "attention" in methods:
            # autoregressive beam search with the left decoder (search.py:251-360); the decoder step runs on the GPU,
            # the beam bookkeeping on the host like the reference's
            def step(hyps):
                return self.engine.decoder_step_topk(encoder_out, encoder_lens, hyps, beam_size, cat_embs, beam_size)
            results["attention"] = attention_beam_search(step, encoder_out.shape[0], encoder_out.shape[1], beam_size,
                                                         self.sos, self.eos, length_penalty)
        if "ctc_greedy_search" in methods:
            results["ctc_greedy_search"] = greedy_results(self.engine.greedy_search(topk_idx, encoder_lens, blank_id))
        if need_beam:
            l2r = r2l = None
            if "attention_rescoring" in methods:
                # fused native path: the n-best stays on the device between the search and the decoder
                toks, tims, olen, scores, nhyp, l2r, r2l = self.engine.beam_search_rescoring(
                    topk_val, topk_idx, encoder_out, encoder_lens, beam_size, blank_id, cat_embs, reverse_weight)
            else:
                toks, tims, olen, scores, nhyp = self.engine.prefix_beam_search_raw(topk_val, topk_idx, encoder_lens,
                                                                                    beam_size, blank_id)
            if "ctc_prefix_beam_search" in methods:
                per_utt = []
                for b in range(toks.shape[0]):
                    n = int(nhyp[b])
                    per_utt.append(([tuple(toks[b, r, :olen[b, r, 0]].tolist()) for r in range(n)],
                                    [float(x) for x in scores[b, :n]],
                                    [tims[b, r, :olen[b, r, 1]].tolist() for r in range(n)]))
                results["ctc_prefix_beam_search"] = prefix_beam_results(per_utt)
            if "attention_rescoring" in methods:
                results["attention_rescoring"] = rescoring_pick_batch(toks, tims, olen, scores, nhyp, l2r, r2l,
                                                                      ctc_weight, reverse_weight)
        return results

    def attention_rescoring(self, prefix_results: List[DecodeResult], encoder_out: torch.Tensor, encoder_lens,
                            ctc_weight: float = 0.0, reverse_weight: float = 0.0, cat_embs=None) -> List[DecodeResult]:
        """asr/wenet/transformer/search.py:363-448, batched over utterances x hypotheses."""
        nbest = [r.nbest for r in prefix_results]
        l2r, r2l = self.engine.rescoring_scores(encoder_out, encoder_lens, nbest, cat_embs, reverse_weight)
        out = []
        for b, r in enumerate(prefix_results):
            out.append(rescoring_pick(r.nbest, r.nbest_scores, r.nbest_times, l2r[b],
                                      None if r2l is None else r2l[b], ctc_weight, reverse_weight))
        return out
