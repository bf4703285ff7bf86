"""N>1 path on CPU: world_size-2 gloo run of the chunk sharding + single all-gather (reverb_b200/dist.py)."""
import os
import pytest
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_and_order():
    from reverb_b200.dist import chunk_plan, sample_range_for_chunks, shard_range
    for n in (0, 1, 2, 7, 120, 121):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= hi - lo <= -(-n // w) for lo, hi in spans)
    # 1 h of 16 kHz audio in 30 s chunks: 359998 frames = 120 full chunks + a 238-frame tail
    m, nc = chunk_plan(3600 * 16000, 2998)
    assert (m, nc) == (359998, 121)
    s0, s1, nfr = sample_range_for_chunks(16, 32, 2998, m)
    assert nfr == 16 * 2998 and s0 == 160 * 16 * 2998 and s1 - s0 == 160 * (nfr - 1) + 400
    p0 = sample_range_for_chunks(0, 16, 2998, m)
    assert p0[1] - s0 == 240                                                  # neighbour overlap


def test_record_roundtrip():
    from reverb_b200.dist import pack_results, unpack_results
    from reverb_b200.search import DecodeResult
    rs = [DecodeResult([5, 6, 7], -12.5, 0.25, [0.5, 0.25, 0.125], [3, 9, 11]),
          DecodeResult((), -0.75, 0.9, [], []),
          DecodeResult([42], -3.0000001, 0.0, None, None)]
    back = unpack_results(pack_results(rs, 5, 8), 8)
    assert len(back) == 3
    for a, b in zip(rs, back):
        assert list(a.tokens) == list(b.tokens) and a.score == b.score and a.confidence == b.confidence
        assert a.times == b.times and a.tokens_confidence == b.tokens_confidence


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from reverb_b200.dist import decode_sharded, shard_range
    from reverb_b200.search import DecodeResult
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_chunks = 7
    calls = []
    def decode_chunks(c0, c1):
        calls.append((c0, c1))
        return [DecodeResult([c, c + 1, 100 + c][: 1 + c %% 3], -float(c) - 0.5, 0.01 * c, [0.5] * (1 + c %% 3),
                             list(range(1 + c %% 3))) for c in range(c0, c1)]
    out = decode_sharded(decode_chunks, n_chunks, 16, torch.device("cpu"))
    c0, c1 = shard_range(n_chunks, rank, world)
    assert calls == ([(c0, c1)] if c1 > c0 else [])        # an empty shard decodes nothing but still gathers
    assert len(out) == n_chunks
    for c, r in enumerate(out):
        assert r.tokens == [c, c + 1, 100 + c][: 1 + c %% 3] and r.score == -float(c) - 0.5
    # the per-step gather of bench.py (RecordGatherer): every rank contributes a fixed block of `per` records
    from reverb_b200.dist import RecordGatherer, unpack_results
    per = 3
    g = RecordGatherer(torch.device("cpu"), per, 16)
    handles = [g.submit([DecodeResult([rank, step, i], -1.0 * i, 0.5, [0.25] * 3, [i, i + 1, i + 2]) for i in range(per)])
               for step in range(2)]
    for step, h in enumerate(handles):
        got = unpack_results(g.wait(h), 16)
        assert len(got) == per * world
        assert [r.tokens for r in got] == [[rk, step, i] for rk in range(world) for i in range(per)]
    if rank == 0:
        print("GATHER_OK", json.dumps([r.tokens for r in out]))
    dist.destroy_process_group()
""")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gloo_gather_over_ranks(tmp_path, world):
    """world 2 is the contract's minimum; 4 and 8 = the scaling run's process counts (8 ranks > the 7 chunks of the
    worker, so one rank owns an EMPTY shard and must still take part in the all-gather)."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GATHER_OK" in r.stdout
