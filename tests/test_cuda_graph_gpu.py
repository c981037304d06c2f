"""GPU: the fast path is CUDA-graph capturable (nothing in micro_step / sync_step / the fused optimiser
synchronises the host or copies from pageable memory when the keep-mask is already on the device), and
a replayed graph reproduces the eager result. This is how the launch-bound shapes (tshirt 1x28x28,
SD latents at B = 1) are meant to be run: one graph launch per optimiser step."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


class TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(5)
        self.c1 = torch.nn.Conv2d(1, 8, 3, padding=1)
        self.c2 = torch.nn.Conv2d(8, 1, 3, padding=1)

    def forward(self, x, timesteps, return_dict=False, **kw):
        return (self.c2(torch.tanh(self.c1(x))),)


def _make(dev, fused_opt):
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.optim import FusedCombineAdamW
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    net = TinyNet().to(dev)
    comb = GradCombiner(net.parameters())
    opt = FusedCombineAdamW(comb, lr=1e-3, betas=(0.95, 0.999), weight_decay=1e-6) if fused_opt else None
    step = UnlearnStep(net, SissDDPMScheduler(), comb, loss_fn="importance_sampling_with_mixture", train_batch_size=32,
                       lambd=0.5, scaling_norm=5.0, max_norm=1.0, inf_guard=True)
    return net, comb, opt, step


@pytest.mark.parametrize("fused_opt", [False, True])
def test_optimizer_step_is_graph_capturable(cuda_device, fused_opt):
    from siss_b200.step import batch_stats
    dev = cuda_device
    torch.backends.cudnn.allow_tf32 = False
    B, shape = 32, (32, 1, 28, 28)                      # delete_tshirt at BASELINE's batch
    g = torch.Generator(device=dev).manual_seed(1)
    static = dict(x0=torch.empty(shape, device=dev), a0=torch.empty(shape, device=dev), noise=torch.empty(shape, device=dev),
                  t=torch.empty(B, dtype=torch.long, device=dev), keep=torch.empty(B, dtype=torch.bool, device=dev))

    def fill(seed):
        gg = torch.Generator(device=dev).manual_seed(seed)
        static["x0"].copy_(torch.rand(shape, device=dev, generator=gg) * 2 - 1)
        static["a0"].copy_(torch.rand(shape, device=dev, generator=gg) * 2 - 1)
        static["noise"].copy_(torch.randn(shape, device=dev, generator=gg))
        static["t"].copy_(torch.randint(0, 1000, (B,), device=dev, generator=gg))
        static["keep"].copy_(torch.rand(B, device=dev, generator=gg) > 0.5)

    def one_step(step, opt):
        out = step.micro_step(static["x0"], static["a0"], static["noise"], static["t"], keep_mask=static["keep"])
        stats = batch_stats(out, 784)
        if opt is None:
            gstats = step.sync_step()
        else:
            step._micro = 0
            gstats = opt.step(scaling_norm=5.0, max_norm=1.0, inf_guard=True)
        return stats, gstats

    # eager reference run (3 optimiser steps on seeds 10, 11, 12)
    net_e, comb_e, opt_e, step_e = _make(dev, fused_opt)
    eager = []
    for seed in (10, 11, 12):
        fill(seed)
        s, gs = one_step(step_e, opt_e)
        eager.append((s.clone(), gs.clone(), torch.cat([p.grad.reshape(-1) for p in net_e.parameters()]).clone(),
                      torch.cat([p.detach().reshape(-1) for p in net_e.parameters()]).clone()))

    # graph run: warm up on a side stream, capture ONE optimiser step, replay it on new data
    net_g, comb_g, opt_g, step_g = _make(dev, fused_opt)
    fill(10)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        one_step(step_g, opt_g)                          # seed 10 executed eagerly = first optimiser step
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    fill(11)
    with torch.cuda.graph(graph):
        s_static, gs_static = one_step(step_g, opt_g)
    # capture does not execute: replay for seed 11, then seed 12
    for i, seed in enumerate((11, 12), start=1):
        fill(seed)
        graph.replay()
        torch.cuda.synchronize()
        torch.testing.assert_close(s_static, eager[i][0], rtol=1e-5, atol=1e-6, equal_nan=True)
        if opt_g is None:
            torch.testing.assert_close(gs_static, eager[i][1], rtol=1e-4, atol=1e-7)
            got = torch.cat([p.grad.reshape(-1) for p in net_g.parameters()])
            torch.testing.assert_close(got, eager[i][2], rtol=2e-4, atol=1e-6)
        else:
            # fused AdamW: the step count lives in device memory and is advanced inside the graph
            # (siss_counter_add), so EVERY replay applies the bias corrections of its own optimiser step.
            got = torch.cat([p.detach().reshape(-1) for p in net_g.parameters()])
            torch.testing.assert_close(got, eager[i][3], rtol=1e-5, atol=1e-6)
            assert int(opt_g.d_step.item()) == i + 1


@pytest.mark.parametrize("loss_fn", ["importance_sampling_with_mixture", "naive_del"])
def test_device_rng_draws_fresh_values_on_every_graph_replay(cuda_device, loss_fn):
    """Opt-in device RNG with the draw index in device memory (DeviceRng(device_counter=...)): a captured optimiser
    step — which draws eps, t and the Bernoulli mask itself — must produce, on replay k, exactly the eager run's
    step with draw index k (not the captured one again)."""
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.rng import DeviceRng
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    dev = cuda_device
    torch.backends.cudnn.allow_tf32 = False
    B, shape = 32, (32, 1, 28, 28)
    g = torch.Generator(device=dev).manual_seed(3)
    x0 = torch.rand(shape, device=dev, generator=g) * 2 - 1
    a0 = torch.rand(shape, device=dev, generator=g) * 2 - 1
    kw = dict(lambd=0.5, scaling_norm=5.0) if loss_fn.startswith("importance") else {}

    def make(device_counter):
        net = TinyNet().to(dev)
        comb = GradCombiner(net.parameters())
        step = UnlearnStep(net, SissDDPMScheduler(), comb, loss_fn=loss_fn, train_batch_size=B, max_norm=1.0,
                           device_rng=DeviceRng(seed=17, device_counter=dev if device_counter else None), **kw)
        return net, step

    def one(step):
        out = step.micro_step(x0, a0)
        return out["timesteps"], step.sync_step()

    net_e, step_e = make(False)
    eager = []
    for _ in range(4):
        ts, gs = one(step_e)
        eager.append((ts.clone(), gs.clone(), torch.cat([p.grad.reshape(-1) for p in net_e.parameters()]).clone()))
    assert not torch.equal(eager[0][0], eager[1][0]) and not torch.equal(eager[1][2], eager[2][2])   # draws do differ

    net_g, step_g = make(True)
    from siss_b200.graph import CapturedStep
    captured = CapturedStep(lambda: one(step_g), warmup=1)     # draw 0 eagerly (side stream), then captured while the
    ts_static, gs_static = captured.outputs                    # device counter reads 1
    graph = captured.graph
    for k in (1, 2, 3):
        graph.replay()
        torch.cuda.synchronize()
        assert int(step_g.device_rng.d_draw.item()) == k + 1
        assert torch.equal(ts_static, eager[k][0]), k
        torch.testing.assert_close(gs_static, eager[k][1], rtol=1e-4, atol=1e-7)
        got = torch.cat([p.grad.reshape(-1) for p in net_g.parameters()])
        torch.testing.assert_close(got, eager[k][2], rtol=2e-4, atol=1e-6)
