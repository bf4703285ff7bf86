#!/usr/bin/env python
"""Benchmark of the hot path: RTFx (audio-seconds / second) of reverb_asr_v1-shaped attention-rescoring
decode of 30 s chunks (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W      # the reference algorithm on the host cores

A "step" = one batch of `--chunks` (default 64) 30 s chunks per GPU through
    fbank -> Conformer encoder -> CTC head -> ctc_prefix_beam_search -> attention_rescoring.
`value`: int16 PCM already resident in HBM when the timed region starts.  `e2e`: the same step through the public
host API with pinned HOST buffers — H2D of the PCM and D2H of the hypotheses inside the timed region.
Weak scaling: every rank decodes its own `--chunks` chunks (chunks are independent units, no data-path collective;
the only collective is the barrier / max-over-ranks of the timing contract).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CHUNK_FRAMES = 2998                     # 30 s of 10 ms frames (snip_edges)
CHUNK_SAMPLES = 480000
METRIC = "rtfx_attention_rescoring_30s_chunks"
UNIT = "audio-seconds/second"


def algorithmic_flops_per_chunk(shape) -> float:
    """Encoder + CTC head FLOPs per 30 s chunk (BASELINE.md §3 formula)."""
    d, ff, L, K, V = shape["d"], shape["ff"], shape["blocks"], shape["kernel"], shape["vocab"]
    T = CHUNK_FRAMES
    T1, F1 = (T - 1) // 2, 39
    Tp, F2 = (T1 - 1) // 2, 19
    mac = T1 * F1 * 9 * d + Tp * F2 * 9 * d * d + Tp * F2 * d * d \
        + L * Tp * (4 * d * ff + 7 * d * d + 3 * Tp * d + K * d) + 2 * Tp * shape["emb_len"] * d * d + Tp * d * V
    return 2.0 * mac


def model_dir_for(shape_name: str) -> str:
    from reverb_b200 import synth
    shape = synth.BENCH_SHAPE if shape_name == "bench" else synth.TEST_SHAPE
    d = os.path.join(os.environ.get("RVB_BENCH_DIR", "/tmp"), f"rvb_bench_model_{shape_name}")
    if not (os.path.exists(os.path.join(d, "synth.pt")) and os.path.exists(os.path.join(d, ".complete"))):
        synth.write_model_dir(d, shape=shape, seed=0, causal=True, cnn_module_norm="layer_norm", reverse_weight=0.3)
        open(os.path.join(d, ".complete"), "w").close()
    return d


def make_pcm(n_chunks: int, seed: int) -> np.ndarray:
    """n_chunks x 30 s of speech-like int16 audio; the first 5 chunks are synthesised, the rest tiled with a
    per-chunk gain so that chunks differ."""
    from reverb_b200 import synth
    base = [synth.synth_audio(30.0, seed=seed + i) for i in range(min(n_chunks, 5))]
    out = np.empty((n_chunks, CHUNK_SAMPLES), dtype=np.int16)
    for i in range(n_chunks):
        g = 1.0 - 0.03 * (i // len(base) % 8)
        out[i] = (base[i % len(base)].astype(np.float32) * g).astype(np.int16)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def pick_threads(orc) -> int:
    """torch CPU throughput is not monotone in the thread count (a 128-way split of a 748-row GEMM thrashes):
    probe the encoder on a 10 s chunk with a few counts and keep the fastest — 'all the threads it can USE'."""
    avail = host_threads()
    cands = sorted({c for c in (avail, 64, 32, 16, 8) if 1 <= c <= avail}, reverse=True)
    feats = torch.randn(1, 998, 80) * 3 + 10
    lens = torch.tensor([998], dtype=torch.int32)
    cat = torch.tensor([1.0, 0.0])
    best, best_t = cands[-1], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        orc.forward_encoder(feats, lens, cat)
        t0 = time.perf_counter()
        orc.forward_encoder(feats, lens, cat)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def run_cpu_reference(model_dir: str, n_chunks: int, steps: int, warmup: int, threads: int):
    """The reference algorithm (oracle port: same ATen CPU operators as the reference's torch.nn graph, same Python
    searches) on the host cores.  One step = `n_chunks` 30 s chunks, batch_size 1 like the reference default.
    Returns (RTFx, seconds per step, threads used)."""
    from oracle import pipeline_ref
    orc = pipeline_ref.OracleASR(model_dir)
    threads = pick_threads(orc) if threads <= 0 else threads
    torch.set_num_threads(threads)
    pcm = make_pcm(n_chunks, seed=4321)
    cat = torch.tensor([1.0, 0.0])
    from oracle import fbank_np

    def step():
        for c in range(n_chunks):
            feats = torch.from_numpy(fbank_np.fbank(pcm[c].astype(np.float32))).unsqueeze(0)
            lens = torch.tensor([feats.shape[1]], dtype=torch.int32)
            orc.decode(["attention_rescoring"], feats, lens, 10, ctc_weight=0.1, reverse_weight=0.0, cat_embs=cat)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_chunks * 30.0 * steps / dt, dt / steps, threads


def _claim_stdout():
    """Only the JSON line may reach stdout: libraries (NCCL prints its version there) are redirected to stderr."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_fd: int, line: dict):
    sys.stdout.flush()
    os.write(real_fd, (json.dumps(line) + "\n").encode())


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rvb", choices=["rvb", "reference"])
    ap.add_argument("--chunks", type=int, default=64, help="30 s chunks per GPU per step")
    ap.add_argument("--shape", default="bench", choices=["bench", "test"])
    ap.add_argument("--reverse_weight", type=float, default=0.0)
    ap.add_argument("--mode", default="attention_rescoring", choices=["attention_rescoring", "ctc_prefix_beam_search"],
                    help="decode mode of the step: the metric's attention_rescoring (default; a superset of BASELINE "
                         "configs[1]) or configs[1] exactly (encoder + ctc_prefix_beam_search)")
    ap.add_argument("--cpu-chunks", type=int, default=1, help="30 s chunks per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=1, help="concurrent decoding lanes (streams + host threads) per GPU")
    ap.add_argument("--profile-step", action="store_true",
                    help="after warm-up run ONE step between cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`)")
    ap.add_argument("--breakdown", action="store_true", help="print a per-stage wall-clock split (synchronised) to stderr")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from reverb_b200 import synth
    shape = synth.BENCH_SHAPE if args.shape == "bench" else synth.TEST_SHAPE
    stages = "fbank+ConformerEncoder+ctc_prefix_beam_search" + ("+attention_rescoring" if args.mode == "attention_rescoring" else "")
    config = {"workload": f"BASELINE configs[1]: {args.chunks}x30s chunks per GPU, {stages}, synthetic reverb_asr_v1 shape "
                          f"(d={shape['d']}, L={shape['blocks']}, V={shape['vocab']})", "mode": args.mode,
              "chunk_frames": CHUNK_FRAMES, "chunks_per_gpu": args.chunks, "beam_size": 10, "ctc_weight": 0.1,
              "reverse_weight": args.reverse_weight, "parallelism": f"chunk-sharded x{world}", "lanes_per_gpu": args.lanes,
              "l2_policy": "inputs larger than L2 (61 MB PCM, multi-GB activations per step); no explicit flush"}

    # ------------------------------------------------------------------ reference arm (host cores)
    if args.impl == "reference":
        if rank != 0:
            return
        mdir = model_dir_for(args.shape)
        val, sec, threads = run_cpu_reference(mdir, args.cpu_chunks, max(args.steps, 1), max(args.warmup, 0), 0)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{args.cpu_chunks} x 30 s chunks per step, batch_size 1, torch "
                                           f"{torch.__version__} CPU fp32"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _emit(real_stdout, line)
        return

    # ------------------------------------------------------------------ CUDA arm
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        mdir = model_dir_for(args.shape)
    if world > 1:
        dist.barrier()
    mdir = model_dir_for(args.shape)
    import reverb_b200
    from reverb_b200 import _lib
    from reverb_b200.engine import launch_count
    asr = reverb_b200.ReverbASR(os.path.join(mdir, "config.yaml"), os.path.join(mdir, "synth.pt"), gpu=local_rank)
    lib = _lib.load()
    eng, model = asr.engine, asr.model
    pcm_host = torch.from_numpy(make_pcm(args.chunks, seed=1234 + 17 * rank)).pin_memory()
    pcm_dev = pcm_host.to(dev)
    cat = torch.tensor([1.0, 0.0])
    lens = torch.full((args.chunks,), CHUNK_FRAMES, dtype=torch.int32)
    stats = {"d2h": 0, "tokens": 0}

    def decode_device(pcm: torch.Tensor):
        feats = eng.fbank_batch(pcm)                                                      # (B, 2998, 80)
        res = model.decode([args.mode], feats, lens, 10, ctc_weight=0.1,
                           reverse_weight=args.reverse_weight, blank_id=asr.blank_id, cat_embs=cat)
        return res[args.mode]

    lanes = None
    if args.lanes > 1:
        from reverb_b200.pipeline import Lanes
        lanes = Lanes(asr, args.lanes)
        assert args.chunks % args.lanes == 0
    per = args.chunks // max(args.lanes, 1)
    lens_lane = lens[:per]

    def lane_job(mdl, pcm):     # pcm: (per, samples) int16, device or pinned host
        if not pcm.is_cuda:
            pcm = pcm.to(dev, non_blocking=True)
        feats = mdl.engine.fbank_batch(pcm)
        res = mdl.decode([args.mode], feats, lens_lane, 10, ctc_weight=0.1,
                         reverse_weight=args.reverse_weight, blank_id=asr.blank_id, cat_embs=cat)
        return res[args.mode]

    def step_resident():
        if lanes is None:
            return decode_device(pcm_dev)
        outs = lanes.run([pcm_dev[i * per:(i + 1) * per] for i in range(args.lanes)], lane_job)
        return [h for o in outs for h in o]

    def step_e2e():
        if lanes is None:
            return decode_device(pcm_host.to(dev, non_blocking=True))
        outs = lanes.run([pcm_host[i * per:(i + 1) * per] for i in range(args.lanes)], lane_job)
        return [h for o in outs for h in o]

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    for _ in range(max(args.warmup, 3)):
        step_resident()

    if args.profile_step:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        print("profiled one step", file=sys.stderr)
        return

    if args.breakdown and rank == 0:
        from reverb_b200.search import rescoring_pick_batch

        def tick(label, fn, acc):
            fn()                                  # first call may allocate; time the second
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize(dev)
            acc.append((label, (time.perf_counter() - t0) * 1e3))
            return out
        acc = []
        feats = tick("fbank", lambda: eng.fbank_batch(pcm_dev), acc)
        enc, enc_lens = tick("encoder", lambda: eng.forward_encoder(feats, lens.numpy(), cat), acc)
        tv, ti, _ = tick("ctc_head+topk", lambda: eng.ctc_topk(enc, 10, 0.0, asr.blank_id), acc)
        tick("prefix_beam only (gpu+copy)", lambda: eng.prefix_beam_search_raw(tv, ti, enc_lens, 10, asr.blank_id), acc)
        raw = tick("prefix_beam+rescoring decoder (fused native call)",
                   lambda: eng.beam_search_rescoring(tv, ti, enc, enc_lens, 10, asr.blank_id, cat, args.reverse_weight), acc)
        tick("host pick", lambda: rescoring_pick_batch(*raw[:5], raw[5], raw[6], 0.1, args.reverse_weight), acc)
        print("BREAKDOWN " + json.dumps({k: round(v, 2) for k, v in acc}), file=sys.stderr)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = launch_count()
    lib.rvb_gemm_profile_begin()
    ms, hyps = timed(step_resident, args.steps)
    gms, gfl, gn = C.c_double(), C.c_double(), C.c_longlong()
    lib.rvb_gemm_profile_end(C.byref(gms), C.byref(gfl), C.byref(gn))
    launches = launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    audio_s = args.chunks * 30.0 * args.steps * world
    value = audio_s / (ms / 1e3)

    # end-to-end through the host API (pinned host PCM in, host hypotheses out)
    step_e2e()
    ms_e2e, hyps = timed(step_e2e, args.steps)
    e2e_val = audio_s / (ms_e2e / 1e3)
    n_tok = sum(len(h.tokens) for h in hyps)
    d2h = int(getattr(eng, "last_d2h_bytes", 0))     # counted by the engine from the arrays the native call fills

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    achieved_tf = (gfl.value / (gms.value / 1e3)) / 1e12 if gms.value > 0 else 0.0
    # dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch, from the committed ncu pass over one step of this
    # same command (profiles/gemm_traffic.json, written from the ncu csv by tools/summarize_dram.py); None if absent
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("shape") == args.shape and tj.get("chunks") == args.chunks:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj.get("source")
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": config,
        "clocks": clk,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(pcm_host.numel() * 2),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tc2_kernel (2-CTA tcgen05 + TMA, all dense layers incl. conv2 implicit GEMM)",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "peak_source": peak_src, "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu)",
                     "traffic_source": traffic_src,
                     "launches_timed": int(gn.value), "kernel_ms_per_step": gms.value / args.steps,
                     "kernel_share_of_step": gms.value / ms if ms > 0 else None,
                     "algorithmic_flops_per_step": gfl.value / args.steps},
        "tokens_per_step": n_tok,
        "encoder_ctc_tflop_per_step": algorithmic_flops_per_chunk(shape) * args.chunks / 1e12,
    }
    if not args.no_cpu_baseline and world == 1:
        t0 = time.time()
        val, sec, threads = run_cpu_reference(mdir, args.cpu_chunks, 1, 0, 0)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{args.cpu_chunks} x 30 s chunks (one timed pass after a thread-count probe), batch_size 1, "
                                          f"oracle port of the reference on torch {torch.__version__} CPU fp32, "
                                          f"{time.time() - t0:.0f} s wall"}
    _emit(real_stdout, line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
