"""N > 1 path of bench.py on CPU: two gloo ranks.  The hot path has no collective (replicas only); what spans
ranks is the timing contract -- max over ranks of the device time, sum over ranks of the audio produced."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank r took (1000 + 500 r) ms to produce 60 s of audio, and (2000 - 300 r) ms end to end for 61 s
        agg = bench.aggregate_ranks(1000.0 + 500.0 * rank, 2000.0 - 300.0 * rank, 60.0, 61.0, world, "cpu")
        out[rank] = agg
    finally:
        dist.destroy_process_group()


def test_two_rank_aggregation_is_max_time_sum_audio():
    world, port = 2, 29500 + os.getpid() % 500
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    for r in (0, 1):
        assert res[r]["ms_max"] == pytest.approx(1500.0)            # slowest rank
        assert res[r]["e2e_ms_max"] == pytest.approx(2000.0)
        assert res[r]["value"] == pytest.approx(2 * 60.0 / 1.5)     # whole-job audio / max time
        assert res[r]["e2e_value"] == pytest.approx(2 * 61.0 / 2.0)


def test_single_rank_is_identity():
    sys.path.insert(0, ROOT)
    import bench
    agg = bench.aggregate_ranks(1234.0, 2345.0, 50.0, 51.0, 1, "cpu")
    assert agg["value"] == pytest.approx(50.0 / 1.234) and agg["e2e_value"] == pytest.approx(51.0 / 2.345)
