"""Concurrent decoding lanes: N host threads, each with its own CUDA stream and its own plan workspace
(`Engine.fork()` — packed weights are shared), so that the latency-bound stages of one batch (CTC prefix beam
search: one CTA per utterance; host-side n-best packing; D2H syncs) overlap with the tensor-core-bound encoder of
another batch.  Chunks are independent units (asr/wenet/cli/reverb.py:214-234), so results are identical to the
sequential schedule; order is preserved.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Sequence

import torch

from .asr_model import ASRModel


class Lanes:
    def __init__(self, asr, n_lanes: int = 2):
        assert n_lanes >= 1
        self.asr = asr
        self.device = asr.device
        self.models: List[ASRModel] = [asr.model]
        for _ in range(n_lanes - 1):
            self.models.append(ASRModel(asr.engine.fork(), asr.configs, asr.configs["output_dim"]))
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_lanes)]
        # one worker PER LANE: a lane's plan (workspace, pinned staging, CUDA stream) is only ever driven by the one
        # host thread that owns it, whatever the number of jobs
        self.pool = ThreadPoolExecutor(max_workers=n_lanes)

    def __len__(self):
        return len(self.models)

    def run(self, jobs: Sequence, fn: Callable):
        """fn(model, job) for every job; lane l processes jobs l, l + n, l + 2n, ... sequentially on its own stream
        and plan; returns results in job order.  Work already queued on the caller's current stream is visible to
        the lanes."""
        jobs = list(jobs)
        n = len(self.models)
        producer = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(producer)
        results = [None] * len(jobs)

        def lane_task(lane):
            torch.cuda.set_device(self.device)
            stream = self.streams[lane]
            with torch.cuda.stream(stream):
                stream.wait_event(ready)
                for i in range(lane, len(jobs), n):
                    results[i] = fn(self.models[lane], jobs[i])
                stream.synchronize()

        futures = [self.pool.submit(lane_task, lane) for lane in range(min(n, len(jobs)))]
        for f in futures:
            f.result()
        return results

    def close(self):
        self.pool.shutdown(wait=True)
