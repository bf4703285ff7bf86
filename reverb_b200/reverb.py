"""`ReverbASR` / `load_model` — the public Python API, a drop-in for the reference's
asr/wenet/cli/reverb.py (same class, method names, argument names, defaults and error
behaviour), running on the native B200 engine.

Differences that are deliberate and documented in DESIGN.md:
  * the model always runs on a CUDA device (`gpu < 0` selects the current device; there
    is no CPU path), whereas the reference's `load_model()` is CPU-only (reverb.py:354-357);
  * RIFF/WAVE audio (PCM 8/16/24/32 bit, float, extensible, multi-channel) is parsed by
    reverb_b200/audio_io.py with torchaudio.load(normalize=False) value conventions, other
    containers go through torchaudio when it has a decoder backend; a file that is not
    16 kHz is resampled on the GPU with torchaudio.transforms.Resample's algorithm
    (csrc/resample.cu + resample.py), where the reference calls torchaudio on the CPU.
"""
from __future__ import annotations

import logging
import shutil
from functools import partial
from itertools import chain
from math import ceil
from pathlib import Path
from typing import Dict, Generator, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F
import yaml

from .asr_model import ASRModel
from .ctc_align import adjust_model_time_offset, ctc_align, hyps_to_ctm, hyps_to_txt
from .engine import Engine, check_beam_size
from .search import DecodeResult
from .text import get_blank_id, init_tokenizer

_FRAME_DOWNSAMPLING_FACTOR = {"linear": 1, "conv2d": 4, "conv2d6": 6, "conv2d8": 8}
CACHED_MODELS_DIR = Path.home() / ".cache/reverb"
_MODELS = {"reverb_asr_v1": "https://huggingface.co/Revai/reverb-asr"}


def _read_wav(path: str) -> Tuple[np.ndarray, int]:
    """samples (channels, n) + sample rate — the `torchaudio.load(normalize=False)` contract (audio_io.py)."""
    from .audio_io import load_audio
    return load_audio(path)


def _load_state_dict(checkpoint: str) -> Dict[str, torch.Tensor]:
    """utils/checkpoint.py:29-80: flat state_dict, optional {'model0': sd} wrapper."""
    sd = torch.load(checkpoint, map_location="cpu", mmap=False)
    if isinstance(sd, dict) and "model0" in sd:
        sd = sd["model0"]
    return sd


class ReverbASR:
    def __init__(self, config, checkpoint, cmvn_path: str | None = None, tokenizer_symbols: str | None = None,
                 bpe_path: str | None = None, gpu: int = -1, overwrite_cmvn: bool = False,
                 precision: str | None = None):
        self.jit = False
        if not torch.cuda.is_available():
            raise RuntimeError("reverb_b200.ReverbASR needs a CUDA device (B200, sm_100a); no CPU fallback exists")
        self.device = torch.device("cuda", gpu if gpu >= 0 else torch.cuda.current_device())
        self.checkpoint = checkpoint
        with open(config, "r") as fin:
            self.configs = yaml.load(fin, Loader=yaml.FullLoader)
        self.configs["cmvn_conf"]["cmvn_file"] = self._make_path_absolute(
            self.configs["cmvn_conf"]["cmvn_file"], cmvn_path)
        self.configs["tokenizer_conf"]["symbol_table_path"] = self._make_path_absolute(
            self.configs["tokenizer_conf"]["symbol_table_path"], tokenizer_symbols)
        self.configs["tokenizer_conf"]["bpe_path"] = self._make_path_absolute(
            self.configs["tokenizer_conf"]["bpe_path"], bpe_path)
        self.tokenizer = init_tokenizer(self.configs)
        self.blank_id = get_blank_id(self.configs, self.tokenizer.symbol_table)
        self.configs["output_dim"] = len(self.tokenizer.symbol_table)

        sd = _load_state_dict(checkpoint)
        # utils/init_model.py:102-117 + load_checkpoint: GlobalCMVN exists only when `cmvn: global_cmvn`; it is built
        # from the stats file and then overwritten by the checkpoint's buffers when the checkpoint holds them.
        # cli/reverb.py:80-85: `overwrite_cmvn` puts the file's stats back (the reference reads the top-level
        # `cmvn_file` / `is_json_cmvn` keys there; the cmvn_conf entries are accepted as well).
        input_dim = self.configs.get("input_dim", 80)
        if self.configs.get("cmvn", None) == "global_cmvn":
            from .cmvn import load_cmvn
            have = "encoder.global_cmvn.mean" in sd and "encoder.global_cmvn.istd" in sd
            ow_file = self.configs.get("cmvn_file", self.configs["cmvn_conf"]["cmvn_file"]) if overwrite_cmvn else None
            if ow_file is not None:
                mean, istd = load_cmvn(ow_file, self.configs.get("is_json_cmvn", self.configs["cmvn_conf"]["is_json_cmvn"]))
                have = False
            elif not have:
                mean, istd = load_cmvn(self.configs["cmvn_conf"]["cmvn_file"], self.configs["cmvn_conf"]["is_json_cmvn"])
            if not have:
                sd["encoder.global_cmvn.mean"] = torch.from_numpy(np.asarray(mean)).float()
                sd["encoder.global_cmvn.istd"] = torch.from_numpy(np.asarray(istd)).float()
        else:
            # no GlobalCMVN module in the reference model: the engine's fused (x - mean) * istd becomes the identity
            sd["encoder.global_cmvn.mean"] = torch.zeros(input_dim)
            sd["encoder.global_cmvn.istd"] = torch.ones(input_dim)
        # precision: 'bf16' (default, throughput) or 'fp32' (bf16x3 tcgen05 passes + fp32 attention: reference-level
        # accuracy, engine.resolve_precision); None -> $RVB_PRECISION.  Not a reference argument: an extension.
        self.engine = Engine(self.configs, sd, self.configs["output_dim"], self.device, precision)
        self.model = ASRModel(self.engine, self.configs, self.configs["output_dim"])
        self.test_conf = self.configs["dataset_conf"]
        self.input_frame_length = self.test_conf["fbank_conf"]["frame_shift"]
        self.output_frame_length = self.input_frame_length * _FRAME_DOWNSAMPLING_FACTOR.get(
            self.configs["encoder_conf"]["input_layer"], 4)
        self._lanes = None

    def set_lanes(self, n_lanes: int):
        """Decode consecutive batches on `n_lanes` concurrent streams / host threads (reverb_b200/pipeline.py).
        1 (default) = the reference's strictly sequential batch loop.  Results do not depend on this setting."""
        from .pipeline import Lanes
        if self._lanes is not None:
            self._lanes.close()
        self._lanes = Lanes(self, n_lanes) if n_lanes > 1 else None

    def _make_path_absolute(self, config_path: str, alternate_path: str | None = None) -> str:
        if alternate_path:
            return alternate_path
        p = Path(config_path)
        if not p.is_absolute():
            p = Path(self.checkpoint).parent / p   # adjacent to the checkpoint
        return p.as_posix()

    # ---------------------------------------------------------------------------------------------
    def compute_feats(self, audio_file: str, resample_rate: int = 16000, num_mel_bins=23, frame_length=25,
                      frame_shift=10, dither=0.0) -> torch.Tensor:
        """(1, m, num_mel_bins) float32 on the device; kernel: csrc/fbank.cu."""
        if num_mel_bins != 80 or frame_length != 25 or frame_shift != 10 or dither != 0.0 or resample_rate != 16000:
            raise NotImplementedError("reverb_b200 fbank kernel is built for 80 bins / 25 ms / 10 ms / no dither @16 kHz")
        pcm, sample_rate = _read_wav(audio_file)
        logging.info(f"detected sample rate: {sample_rate}")
        ch0 = np.array(pcm[0], copy=True)                      # channel 0 (kaldi.fbank channel=-1 -> 0)
        if ch0.dtype != np.int16:
            # `waveform.to(torch.float)` of the reference (cli/reverb.py:124): the sample VALUES as they are (uint8 /
            # int32 / float32) — the kernels take int16 or float32 input
            ch0 = ch0.astype(np.float32)
        wave_dev = torch.from_numpy(ch0).pin_memory().to(self.device, non_blocking=True)
        if sample_rate != resample_rate:
            # torchaudio.transforms.Resample on channel 0 (the reference resamples every channel, then keeps the first)
            wave_dev = self.engine.resample(wave_dev, sample_rate, resample_rate)
        if wave_dev.numel() < 400:
            raise AssertionError(f"choose a window size 400 that is [2, {wave_dev.numel()}]")  # torchaudio's check
        return self.engine.fbank(wave_dev).unsqueeze(0)

    def feats_batcher(self, infeats: torch.Tensor, chunk_size: int, batch_size: int
                      ) -> Generator[Tuple[torch.Tensor, torch.Tensor], None, None]:
        """Fixed-length chunks with no overlap; only the final chunk is zero-padded (in feature space)."""
        nbins = self.test_conf["fbank_conf"]["num_mel_bins"]
        per_batch = chunk_size * batch_size
        num_batches = ceil(infeats.shape[1] / per_batch)
        for b in range(num_batches):
            fb = infeats[:, b * per_batch:(b + 1) * per_batch, :]
            lens = torch.tensor([chunk_size] * batch_size, dtype=torch.int32)
            if b == num_batches - 1:
                last = ceil(fb.shape[1] / chunk_size)
                lens = torch.tensor([chunk_size] * last, dtype=torch.int32)
                pad = chunk_size * last - fb.shape[1]
                if pad > 0:
                    lens[-1] -= pad
                    fb = F.pad(fb, (0, 0, 0, pad, 0, 0), mode="constant", value=0)
            yield fb.reshape(-1, chunk_size, nbins), lens

    def transcribe_modes(self, audio_file, modes: List[str], format: str = "txt", verbatimicity: float = 1.0,
                         chunk_size: int = 2051, batch_size: int = 1, beam_size: int = 10,
                         decoding_chunk_size: int = -1, num_decoding_left_chunks: int = -1, ctc_weight: float = 0.1,
                         simulate_streaming: bool = False, reverse_weight: float = 0.0, blank_penalty: float = 0.0,
                         length_penalty: float = 0.0, timings_adjustment: float = 230) -> list[str]:
        check_beam_size(beam_size)        # fail before any audio is read / decoded (limit: engine.MAX_BEAM_SIZE)
        fc = self.test_conf["fbank_conf"]
        feats = self.compute_feats(audio_file, num_mel_bins=fc["num_mel_bins"], frame_length=fc["frame_length"],
                                   frame_shift=fc["frame_shift"])
        with torch.no_grad():
            cat_embs = torch.tensor([verbatimicity, 1.0 - verbatimicity])

            kw = dict(decoding_chunk_size=decoding_chunk_size, num_decoding_left_chunks=num_decoding_left_chunks,
                      ctc_weight=ctc_weight, simulate_streaming=simulate_streaming, reverse_weight=reverse_weight,
                      context_graph=None, blank_id=self.blank_id, blank_penalty=blank_penalty,
                      length_penalty=length_penalty, infos={"tasks": ["transcribe"], "langs": ["en"]}, cat_embs=cat_embs)

            def decode_batch(model, batch):
                feats_batch, feats_lengths = batch
                return model.decode(modes, feats_batch, feats_lengths, beam_size, **kw)

            batches = self.feats_batcher(feats, chunk_size, batch_size)
            if self._lanes is not None:
                results = self._lanes.run(list(batches), decode_batch)
            else:
                # the reference's sequential batch loop (cli/reverb.py:214-234), software-pipelined on one stream
                results = list(self.model.decode_stream(batches, modes, beam_size, **kw))
        return [get_output(format, self.tokenizer, Path(audio_file).name,
                           list(chain(*(hyp[mode] for hyp in results))), timings_adjustment, chunk_size,
                           self.input_frame_length, self.output_frame_length) for mode in modes]

    def transcribe(self, audio_file, mode: str = "ctc_prefix_beam_search", format: str = "txt",
                   verbatimicity: float = 1.0, chunk_size: int = 2051, batch_size: int = 1, beam_size: int = 10,
                   decoding_chunk_size: int = -1, num_decoding_left_chunks: int = -1, ctc_weight: float = 0.1,
                   simulate_streaming: bool = False, reverse_weight: float = 0.0, blank_penalty: float = 0.0,
                   length_penalty: float = 0.0, timings_adjustment: float = 230) -> str:
        return self.transcribe_modes(
            audio_file, modes=[mode], format=format, verbatimicity=verbatimicity, chunk_size=chunk_size,
            batch_size=batch_size, beam_size=beam_size, decoding_chunk_size=decoding_chunk_size,
            num_decoding_left_chunks=num_decoding_left_chunks, ctc_weight=ctc_weight,
            simulate_streaming=simulate_streaming, reverse_weight=reverse_weight, blank_penalty=blank_penalty,
            length_penalty=length_penalty, timings_adjustment=timings_adjustment)[0]


def get_output(format: str, tokenizer, audio_name: str, hyps: List[DecodeResult], timings_adjustment_ms: int,
               chunk_size: int, input_frame_length: int, output_frame_length: int) -> str:
    """One hypothesis per chunk -> words -> CTM lines / text (reference: cli/reverb.py:292-321)."""
    if format == "txt":
        render, delimiter = hyps_to_txt, " "
    elif format == "ctm":
        render, delimiter = partial(hyps_to_ctm, audio_name), "\n"
    else:
        raise ValueError("Invalid output format.")
    lines: List[str] = []
    time_shift_ms = 0
    for hyp in hyps:
        words = ctc_align(hyp.tokens, hyp.times, hyp.tokens_confidence, tokenizer, output_frame_length, time_shift_ms)
        words = adjust_model_time_offset(words, timings_adjustment_ms)
        time_shift_ms += chunk_size * input_frame_length
        lines.extend(render(words))
    return delimiter.join(lines)


def load_model(model: str, gpu: int = -1, precision: str | None = None) -> ReverbASR:
    """Loads a reverb model from a directory (config.yaml + first *.pt) or by pretrained name."""
    if Path(model).exists():
        model_dir = Path(model)
        config_path = model_dir / "config.yaml"
        checkpoint_path = list(model_dir.glob("*.pt"))[0]
    elif model in _MODELS:
        model_dir = CACHED_MODELS_DIR / model
        config_path = model_dir / "config.yaml"
        checkpoint_path = model_dir / f"{model}.pt"
        if not (CACHED_MODELS_DIR.exists() and model_dir.exists() and config_path.exists()
                and checkpoint_path.exists()):
            CACHED_MODELS_DIR.parent.mkdir(exist_ok=True, parents=True)
            shutil.rmtree(model_dir, ignore_errors=True)
            download_model(_MODELS[model], model_dir)
    else:
        raise ValueError("Please specify a local path to a model or one of our pretrained models: "
                         f"{','.join(get_available_models())}")
    config_path, checkpoint_path = config_path.resolve(), checkpoint_path.resolve()
    logging.info(f"Loading the model with {config_path = } and {checkpoint_path = }")
    return ReverbASR(str(config_path), str(checkpoint_path), gpu=gpu, precision=precision)


def get_available_models():
    return list(_MODELS.keys())


def download_model(url: str, root: str):
    """Clones the model repository at `url` into `root` (needs network + GitPython)."""
    from git import Repo
    Repo.clone_from(url, root)
