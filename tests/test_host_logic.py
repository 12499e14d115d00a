"""Host-side logic that runs without a GPU: synthetic corpus generator, stream sharding, and the N>1 span gather
(world_size 2 on gloo)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orz_b200 import dist as odist
from tools import textgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_markov_text_is_deterministic_and_text_shaped():
    a = textgen.markov_text(200_000)
    b = textgen.markov_text(200_000)
    assert (a == b).all()
    assert textgen.markov_text(1000, seed=1).tobytes() != a[:1000].tobytes()
    allowed = set([9, 10] + list(range(0x20, 0x7F)))
    assert set(np.unique(a).tolist()) <= allowed
    assert 0.12 < (a == 0x20).mean() < 0.25  # spaces about as frequent as in enwik8


def test_shard_streams_partition():
    for n in (0, 1, 7, 80):
        for world in (1, 2, 8):
            parts = [odist.shard_streams(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n_local = 3 if rank == 0 else 1
    spans = [torch.from_numpy(rng.integers(0, 256, size=int(rng.integers(1, 5000)), dtype=np.uint8)) for _ in range(n_local)]
    if rank == 1:
        spans.append(torch.empty(0, dtype=torch.uint8))  # an empty stream's span
    lengths, gathered = odist.gather_spans(spans, torch.device("cpu"))
    ok = lengths[rank] == [s.numel() for s in spans] and torch.equal(gathered[rank], torch.cat(spans))
    q.put((rank, ok, [len(l) for l in lengths], [int(g.numel()) for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_spans_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2] == res[1][2] == [3, 2]  # both ranks see both length tables
    assert res[0][3] == res[1][3]


def test_bench_reference_arm_cli_runs_on_cpu():
    # the reference arm needs no GPU: tiny debug corpus, one step
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--bytes", "300000", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0
