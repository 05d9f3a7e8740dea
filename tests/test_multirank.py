"""CPU: the N > 1 host logic with world_size 2 on the gloo backend - stream partition, per-rank seeds, the timing
protocol's max-over-ranks / sum-of-units reduction - and the reference arm under a 2-process launch (rank 0 alone
works and prints one JSON line)."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ffvvc_b200 import streams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = streams.streams_of_rank(8, rank, world)
        units = 1000.0 * len(mine)
        elapsed = 10.0 + 5.0 * rank                      # rank 1 is the slow one
        dist.barrier()
        total, slowest = streams.aggregate(units, elapsed, dist)
        out[rank] = (mine, streams.seed_of_rank(12345, rank), total, slowest)
    finally:
        dist.destroy_process_group()


def test_partition_and_timing_protocol_world2():
    world, port = 2, 29650 + os.getpid() % 200
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert sorted(res[0][0] + res[1][0]) == list(range(8)) and not set(res[0][0]) & set(res[1][0])
    assert res[0][0] == [0, 2, 4, 6] and res[1][0] == [1, 3, 5, 7]
    assert res[0][1] != res[1][1]                         # different content per rank
    for r in (0, 1):
        assert res[r][2] == 8000.0 and res[r][3] == 15.0  # sum of units, max of times, identical on every rank
    assert streams.throughput_mpix(8000.0e6, 15.0) == pytest.approx(8000.0e6 / 0.015 / 1e6)


def test_single_process_aggregate_is_identity():
    assert streams.aggregate(5.0, 2.0) == (5.0, 2.0)
    with pytest.raises(ValueError):
        streams.streams_of_rank(4, 2, 2)


def test_reference_arm_two_processes_print_one_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        pytest.skip("oracle not built")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29850 + os.getpid() % 100), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--width", "416", "--height", "240", "--cpu-threads", "2"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["cores"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
