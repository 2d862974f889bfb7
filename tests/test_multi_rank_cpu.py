"""N>1 path on the CPU: two gloo ranks shard the clips of one batch (runner.video_gen_sharded logic), run
the AR x diffusion loop on their shard with the lowered program executed by the CPU op interpreter, and
meet in ONE all-gather; the result must equal the single-rank result (per-clip noise streams)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from common import make_module, step_noise
        from mcvd_b200 import detfill, runner
        from mcvd_b200.program import Engine
        from op_interpreter import Interpreter
        torch.set_num_threads(2)
        cfg, net, sd = make_module("tiny", "cpu")
        net._engine = Engine(net, _test_backend=Interpreter())
        n_clips, L = 3, 4                                   # uneven shards: rank 0 gets 2 clips, rank 1 gets 1
        cfg.sampling.subsample = L
        x, cond_all = detfill.synthetic_inputs(cfg, n_clips)
        shape1 = (1,) + tuple(x.shape[1:])

        def run(lo, hi):
            # per-clip init and per-clip, per-step noise keyed by the GLOBAL clip index
            init = lambda i, shape: torch.cat([detfill.normal(f"init{g}_{i}", shape1) for g in range(lo, hi)])
            noise = lambda i: [torch.cat([detfill.normal(f"z{g}_{i}_{s}", shape1) for g in range(lo, hi)])
                               for s in range(L - 1)]
            return runner.video_gen_clips(cfg, net, cond_all[lo:hi], 5, init_fn=init, noise_fn=noise)

        lo, hi = runner.shard_range(n_clips, rank, world)
        local = run(lo, hi)
        full = runner.gather_clips(local, n_clips, rank, world)     # the one collective of the path
        if rank == 0:
            single = run(0, n_clips)
            q.put((tuple(full.shape), bool(torch.equal(full, single)), float((full - single).abs().max())))
    finally:
        dist.destroy_process_group()


def test_two_rank_clip_sharding_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(600) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    shape, equal, err = q.get(timeout=10)
    assert shape == (3, 5, 32, 32)
    # The CPU interpreter runs oneDNN convolutions whose rounding depends on the batch size, so equality here
    # is to fp32 noise; the CUDA kernels are batch-composition invariant and tests/test_gpu_model.py asserts
    # torch.equal for them (test_full_size_properties_cfg2).
    assert err < 1e-3, f"sharded result differs from single-rank result (max abs {err})"
