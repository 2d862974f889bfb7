"""Host-side lowering (mcvd_b200/program.py, samplers.py, runner.py) checked on the CPU: the op program
is executed by tests/op_interpreter.py and compared with the oracle and the golden fixtures."""
import numpy as np
import pytest
import torch

from common import golden, make_module, max_err, step_noise
from mcvd_b200 import detfill, samplers, runner
from mcvd_b200.program import Engine
from op_interpreter import Interpreter
from oracle import mcvd_oracle as O


def cpu_module(name, conv_mode):
    cfg, net, sd = make_module(name, "cpu")
    eng = Engine(net, _test_backend=Interpreter())
    eng.conv_mode = conv_mode
    net._engine = eng
    return cfg, net, sd


@pytest.mark.parametrize("name", ["tiny", "tiny_spade", "tiny_rgb"])
@pytest.mark.parametrize("conv_mode", ["umma", "simt"])
def test_forward_lowering_matches_oracle_and_golden(name, conv_mode):
    cfg, net, sd = cpu_module(name, conv_mode)
    B = cfg.bench_batch
    x, cond = detfill.synthetic_inputs(cfg, B)
    g = golden(name)
    for t in (0, 37, 990):
        tt = torch.full((B,), t, dtype=torch.long)
        mine = net(x, tt, cond=cond)
        ref = O.unet_forward(cfg, sd, x, tt, cond)
        assert max_err(mine, ref) < 5e-5, (name, t, max_err(mine, ref))
        assert max_err(mine, torch.from_numpy(g[f"eps_t{t}"])) < 5e-5
    P = net.engine().program(B)
    if conv_mode == "umma":
        assert P.n_umma > 0


@pytest.mark.parametrize("name", ["tiny", "tiny_spade"])
def test_samplers_lowering(name):
    cfg, net, sd = cpu_module(name, "umma")
    B = cfg.bench_batch
    L = cfg.sampling.subsample
    x, cond = detfill.synthetic_inputs(cfg, B)
    g = golden(name)
    zs = step_noise(x.shape, L)
    out = samplers.ddpm_sampler(x.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=L,
                                clip_before=True, noise_list=zs)
    assert out.shape == (1,) + tuple(x.shape)
    assert max_err(out[0], torch.from_numpy(g["ddpm"])) < 2e-3
    out = samplers.ddim_sampler(x.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=L,
                                clip_before=True, log=False)
    assert max_err(out[0], torch.from_numpy(g["ddim"])) < 2e-3
    out = samplers.FPNDM_sampler(x.clone(), net, cond=cond, final_only=True, subsample_steps=L, clip_before=True,
                                 log=False)
    assert max_err(out[0], torch.from_numpy(g["fpndm"])) < 2e-3
    # non-final_only returns the trajectory, same length as the reference (L steps + denoise)
    traj = samplers.ddpm_sampler(x.clone(), net, cond=cond, final_only=False, denoise=True, subsample_steps=L,
                                 noise_list=zs)
    assert traj.shape[0] == L + 1
    assert max_err(traj[-1], torch.from_numpy(g["ddpm"])) < 2e-3


def test_video_gen_loop_lowering():
    name = "tiny"
    cfg, net, sd = cpu_module(name, "umma")
    B = cfg.bench_batch
    L = cfg.sampling.subsample
    x, cond = detfill.synthetic_inputs(cfg, B)
    g = golden(name)
    nfp = cfg.sampling.num_frames_pred
    vid = runner.video_gen_clips(cfg, net, cond, nfp,
                                 init_fn=lambda i, shape: detfill.normal(f"ar_init{i}", shape),
                                 noise_fn=lambda i: step_noise(x.shape, L, tag=f"ar{i}_z"))
    ref = torch.from_numpy(g["video"])
    assert vid.shape == ref.shape
    assert O.psnr01(vid, ref) > 50.0
