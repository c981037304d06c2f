"""TEST INFRASTRUCTURE ONLY — numpy restatement of the counter-based generator behind the opt-in device RNG
(`siss_randn`, `siss_draw_rows`, `siss_add_noise_mixture_rng`).

The reference draws noise / timesteps with torch's generators and the Bernoulli mask with CPU `torch.rand`
(delete_celeb.py:581,593; losses/ddpm_deletion_loss.py:18); a device-side counter-based stream is a NEW, opt-in
seed semantic, so there is no reference output to pin. What is pinned instead:
  * Philox4x32-10 (Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3", SC'11) against the
    known-answer vectors published with Random123 (tests/test_philox_cpu.py);
  * the index -> counter mapping and the uint32 -> float transforms, which the CUDA kernels must reproduce
    (integers exactly, Box-Muller outputs to float32 intrinsic accuracy).

Stream definition (shared with csrc/philox.cuh):
  key      = (seed & 0xffffffff, seed >> 32)
  noise    : counter = (c & 0xffffffff, c >> 32, draw & 0xffffffff, draw >> 32), c = global_element_index // 4,
             draw < 2**62; the call's four words give elements 4c .. 4c+3:
             (z0, z1) = box_muller(w0, w1), (z2, z3) = box_muller(w2, w3)
  per row  : counter = (r & 0xffffffff, r >> 32, draw & 0xffffffff, (draw >> 32) | 0x80000000), r = global row index;
             t = t_lo + w0 % (t_hi - t_lo);  keep = uniform(w1) > lambd        (torch.rand(B) > lambd, :18)
  aux      : counter = (c & 0xffffffff, c >> 32, draw & 0xffffffff, (draw >> 32) | 0x40000000): a second element-indexed
             tensor of the same draw (EraseDiff's uniform forget target, losses/ddpm_deletion_loss.py:75);
             element e takes word e % 4 of call e // 4, u = (w >> 8) * 2**-24 in [0, 1)
  uniform(w)    = float32(w) * 2**-32 + 2**-33                  in (0, 1]
  box_muller(a, b): rad = sqrt(-2 ln uniform(a)); ang = 2 pi uniform(b); (rad cos ang, rad sin ang)
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter [..., 4] uint32, key [..., 2] uint32 (broadcastable) -> [..., 4] uint32."""
    c = np.array(counter, dtype=np.uint64, copy=True)
    k = np.broadcast_to(np.asarray(key, dtype=np.uint64), c.shape[:-1] + (2,)).copy()
    for rnd in range(10):
        if rnd:
            k[..., 0] = (k[..., 0] + np.uint64(W0)) & MASK
            k[..., 1] = (k[..., 1] + np.uint64(W1)) & MASK
        p0, p1 = M0 * c[..., 0], M1 * c[..., 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = np.stack([hi1 ^ c[..., 1] ^ k[..., 0], lo1, hi0 ^ c[..., 3] ^ k[..., 1], lo0], axis=-1)
    return c.astype(np.uint32)


def _key(seed: int) -> np.ndarray:
    return np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)


def uniform32(w: np.ndarray) -> np.ndarray:
    return w.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)


def noise_words(n: int, seed: int, draw: int, elem_offset: int = 0) -> np.ndarray:
    """The uint32 word behind each of the n elements [elem_offset, elem_offset + n)."""
    e = np.arange(elem_offset, elem_offset + n, dtype=np.uint64)
    c = e >> np.uint64(2)
    ctr = np.stack([c & MASK, c >> np.uint64(32), np.full_like(c, draw & 0xFFFFFFFF), np.full_like(c, (draw >> 32) & 0x3FFFFFFF)],
                   axis=-1)
    out = philox4x32_10(ctr, _key(seed))
    return out, (e & np.uint64(3)).astype(np.int64)


def randn(n: int, seed: int, draw: int, elem_offset: int = 0) -> np.ndarray:
    """float64 values of the normal stream (Box-Muller evaluated in float64 on the float32 uniforms)."""
    words, lane = noise_words(n, seed, draw, elem_offset)
    pair = lane >> 1
    a = np.take_along_axis(words, (2 * pair)[:, None], axis=1)[:, 0]
    b = np.take_along_axis(words, (2 * pair + 1)[:, None], axis=1)[:, 0]
    u1, u2 = uniform32(a).astype(np.float64), uniform32(b).astype(np.float64)
    rad, ang = np.sqrt(-2.0 * np.log(u1)), 2.0 * np.pi * u2
    return np.where(lane & 1, rad * np.sin(ang), rad * np.cos(ang))


def draw_rows(B: int, seed: int, draw: int, t_lo: int, t_hi: int, lambd: float, row_offset: int = 0):
    """(timesteps int64 [B], keep bool [B]) of global rows [row_offset, row_offset + B)."""
    r = np.arange(row_offset, row_offset + B, dtype=np.uint64)
    ctr = np.stack([r & MASK, r >> np.uint64(32), np.full_like(r, draw & 0xFFFFFFFF),
                    np.full_like(r, ((draw >> 32) & 0x3FFFFFFF) | 0x80000000)], axis=-1)
    w = philox4x32_10(ctr, _key(seed))
    t = t_lo + (w[:, 0].astype(np.int64) % (t_hi - t_lo))
    keep = uniform32(w[:, 1]) > np.float32(lambd)
    return t, keep


def rand_aux(n: int, seed: int, draw: int, elem_offset: int = 0) -> np.ndarray:
    """float32 uniforms in [0, 1) of the aux stream for elements [elem_offset, elem_offset + n) — exact."""
    e = np.arange(elem_offset, elem_offset + n, dtype=np.uint64)
    c = e >> np.uint64(2)
    ctr = np.stack([c & MASK, c >> np.uint64(32), np.full_like(c, draw & 0xFFFFFFFF),
                    np.full_like(c, ((draw >> 32) & 0x3FFFFFFF) | 0x40000000)], axis=-1)
    words = philox4x32_10(ctr, _key(seed))
    w = np.take_along_axis(words, (e & np.uint64(3)).astype(np.int64)[:, None], axis=1)[:, 0]
    return (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
