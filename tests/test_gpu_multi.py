"""2-GPU clip sharding over NCCL: rank r generates its clips with zero communication, one all-gather at
the end; the gathered frames must equal the single-GPU result bit for bit (in-kernel Philox noise keyed by
the global clip id, per-clip initial noise).  Skipped on boxes with fewer than 2 GPUs."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from common import make_module
        from mcvd_b200 import detfill, runner
        cfg, net, sd = make_module("tiny", f"cuda:{rank}")
        cfg.sampling.subsample = 10
        n_clips = 5
        _, cond_all = detfill.synthetic_inputs(cfg, n_clips)
        full = runner.video_gen_sharded(cfg, net, cond_all.to(f"cuda:{rank}"), rank, world, philox_seed=99, init_seed=7,
                                        num_frames_pred=5)
        if rank == 0:
            single = runner.video_gen_sharded(cfg, net, cond_all.to("cuda:0"), 0, 1, philox_seed=99, init_seed=7,
                                              num_frames_pred=5)
            q.put((tuple(full.shape), bool(torch.equal(full, single)), float((full - single).abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharding_bit_exact():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(600) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    shape, equal, err = q.get(timeout=10)
    assert shape == (5, 5, 32, 32)
    assert equal, f"2-GPU result differs from 1-GPU result (max abs {err})"
