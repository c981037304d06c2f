"""CPU property tests (hypothesis) of the host-side bookkeeping that every rank must agree on: row sharding, the global
Bernoulli mask, and the flat dual-buffer layout of GradCombiner (16-byte alignment of every parameter view, padding to a
multiple of 4 * world so every rank's shard is aligned, views that tile the buffer without overlap)."""
import torch
from hypothesis import given, settings, strategies as st

from siss_b200 import parallel
from siss_b200.grad_combine import GradCombiner


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 500), st.integers(1, 16))
def test_shard_bounds_partition_the_batch(B, world):
    spans = [parallel.shard_bounds(B, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == B
    for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
        assert lo <= hi == lo2                               # contiguous, ordered, no gap, no overlap
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)      # remainder goes to the low ranks
    t = torch.arange(B)
    assert torch.equal(torch.cat([parallel.shard_rows(t, r, world) for r in range(world)]), t)


@settings(max_examples=50, deadline=None)
@given(st.integers(1, 64), st.integers(1, 8), st.floats(0.0, 1.0), st.integers(0, 2 ** 31 - 1))
def test_global_keep_mask_is_one_draw_sliced(B, world, lambd, seed):
    torch.manual_seed(seed)
    full = torch.rand(B) > lambd                              # the reference's draw (losses/ddpm_deletion_loss.py:18)
    parts = []
    for r in range(world):
        torch.manual_seed(seed)                               # every rank seeds identically, as under the reference's launcher
        parts.append(parallel.global_keep_mask(B, lambd, r, world))
    assert torch.equal(torch.cat(parts), full)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(1, 70), min_size=1, max_size=8))
def test_flat_buffer_layout(sizes):
    params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
    comb = GradCombiner(params, distributed=False)
    assert comb.num_params == sum(sizes) and comb.total % 4 == 0 and comb.total >= comb.num_params
    end = 0
    for p, off in zip(comb.params, comb.offsets):
        assert off % 4 == 0 and off >= end                    # 16-byte aligned fp32 offset, no overlap with the previous view
        end = off + p.numel()
        assert p.grad is not None and p.grad.data_ptr() == comb.g_x.data_ptr() + 4 * off     # grads start out as views of G_x
    assert end <= comb.total
    comb.begin_a()
    assert all(p.grad.data_ptr() == comb.g_a.data_ptr() + 4 * off for p, off in zip(comb.params, comb.offsets))
    comb.after_backward_a()
    assert all(p.grad.data_ptr() == comb.g_x.data_ptr() + 4 * off for p, off in zip(comb.params, comb.offsets))
