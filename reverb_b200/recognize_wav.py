"""`reverb` command line — same flags, defaults and output layout as the reference's
asr/wenet/bin/recognize_wav.py (flags :33-145, `<result_dir>/<mode>/<audio stem>.ctm` :177-204).

    python -m reverb_b200.recognize_wav --model <dir> --audio_file a.wav --result_dir out
"""
from __future__ import annotations

import argparse
import logging
import os
from pathlib import Path

MODES = ["attention", "ctc_greedy_search", "ctc_prefix_beam_search", "attention_rescoring", "joint_decoding"]


def get_args(argv=None):
    from .reverb import get_available_models
    p = argparse.ArgumentParser(description="Run automatic speech recognition on a given wav file using the Rev model.")
    p.add_argument("--audio_file", required=True, help="Audio to transcribe")
    p.add_argument("--config", default=None, help="Path to config file")
    p.add_argument("--checkpoint", default=None, help="Path to Reverb model checkpoint")
    p.add_argument("--model", default=None,
                   help="Path to directory containing config and checkpoint for a reverb model or the name of a "
                        f"pretrained model from: {','.join(get_available_models())}")
    p.add_argument("--gpu", type=int, default=-1, help="gpu id (this engine always runs on a GPU; -1 = current device)")
    p.add_argument("--tokenizer-symbols", help="Path to tk.units.txt. Overrides the config path.")
    p.add_argument("--bpe-path", help="Path to tk.model. Overrides the config path.")
    p.add_argument("--cmvn-path", help="Path to cmvn. Overrides the config path.")
    p.add_argument("--beam_size", type=int, default=10, help="beam size for search")
    p.add_argument("--length_penalty", type=float, default=0.0,
                   help="length penalty for attention decoding and joint decoding modes")
    p.add_argument("--blank_penalty", type=float, default=0.0, help="blank penalty")
    p.add_argument("--result_dir", required=True, help="asr result file")
    p.add_argument("--batch_size", type=int, default=1, help="Number of chunks that are decoded in parallel")
    p.add_argument("--chunk_size", type=int, default=2051, help="Size of each chunk that is decoded, in frames")
    p.add_argument("--modes", nargs="+", choices=MODES, default=["attention_rescoring"],
                   help="One or more supported decoding mode.")
    p.add_argument("--ctc_weight", type=float, default=0.1, help="ctc weight for attention rescoring decode mode")
    p.add_argument("--decoding_chunk_size", type=int, default=-1,
                   help="decoding chunk size, <0: full chunk (the only mode this engine builds)")
    p.add_argument("--num_decoding_left_chunks", type=int, default=-1, help="number of left chunks for decoding")
    p.add_argument("--simulate_streaming", action="store_true", help="simulate streaming inference")
    p.add_argument("--reverse_weight", type=float, default=0.0,
                   help="right to left weight for attention rescoring decode mode")
    p.add_argument("--overwrite_cmvn", action="store_true",
                   help="overwrite CMVN params in model with those in config file")
    p.add_argument("--verbatimicity", type=float, default=1.0,
                   help="0.0 = nonverbatim ... 1.0 = verbatim; passed to the language-specific layers")
    p.add_argument("--timings_adjustment", type=float, default=230,
                   help="Subtract timings_adjustment milliseconds from each timestamp")
    p.add_argument("--log_level", choices=["DEBUG", "INFO", "WARNING", "ERROR", "CRITICAL"], default="INFO")
    return p.parse_args(argv)


def main(argv=None):
    args = get_args(argv)
    logging.basicConfig(level=args.log_level, format="%(asctime)s %(filename)s %(levelname)s: %(message)s")
    from .reverb import ReverbASR, load_model
    by_name = args.model is not None
    by_files = args.checkpoint is not None and args.config is not None
    if by_name == by_files:
        raise RuntimeError("One of either --model or (--checkpoint and --config) must be set.")
    if by_name:
        asr = load_model(args.model, gpu=args.gpu)
    else:
        asr = ReverbASR(args.config, args.checkpoint, cmvn_path=args.cmvn_path,
                        tokenizer_symbols=args.tokenizer_symbols, bpe_path=args.bpe_path, gpu=args.gpu,
                        overwrite_cmvn=args.overwrite_cmvn)
    targets = {}
    for mode in args.modes:
        out_dir = os.path.join(args.result_dir, mode)
        os.makedirs(out_dir, exist_ok=True)
        targets[mode] = Path(out_dir) / Path(args.audio_file).with_suffix(".ctm").name
    outputs = asr.transcribe_modes(
        args.audio_file, modes=args.modes, format="ctm", verbatimicity=args.verbatimicity,
        chunk_size=args.chunk_size, batch_size=args.batch_size, beam_size=args.beam_size,
        decoding_chunk_size=args.decoding_chunk_size, num_decoding_left_chunks=args.num_decoding_left_chunks,
        ctc_weight=args.ctc_weight, simulate_streaming=args.simulate_streaming, reverse_weight=args.reverse_weight,
        blank_penalty=args.blank_penalty, length_penalty=args.length_penalty,
        timings_adjustment=args.timings_adjustment)
    for mode, text in zip(args.modes, outputs):
        with targets[mode].open(mode="w") as fp:
            fp.write(text)


if __name__ == "__main__":
    main()
