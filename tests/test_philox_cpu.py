"""CPU: pin the counter-based generator's restatement (oracle/philox.py) — Philox4x32-10 against the known-answer
vectors published with Random123 (kat_vectors: philox4x32 10), and the stream's index mapping / statistics."""
import numpy as np

from oracle import philox as P

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        got = P.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(x) for x in got) == want
    # vectorised evaluation agrees with one-at-a-time
    ctrs = np.array([k[0] for k in KAT], dtype=np.uint32)
    keys = np.array([k[1] for k in KAT], dtype=np.uint32)
    assert [tuple(int(x) for x in r) for r in P.philox4x32_10(ctrs, keys)] == [k[2] for k in KAT]


def test_stream_is_a_function_of_the_global_index():
    full = P.randn(1000, seed=42, draw=3)
    assert np.array_equal(P.randn(300, seed=42, draw=3, elem_offset=437), full[437:737])   # unaligned slice
    assert not np.array_equal(P.randn(1000, seed=42, draw=4), full)                         # draws are independent
    assert not np.array_equal(P.randn(1000, seed=43, draw=3), full)
    t, k = P.draw_rows(64, seed=42, draw=3, t_lo=0, t_hi=1000, lambd=0.5)
    t2, k2 = P.draw_rows(16, seed=42, draw=3, t_lo=0, t_hi=1000, lambd=0.5, row_offset=32)
    assert np.array_equal(t2, t[32:48]) and np.array_equal(k2, k[32:48])


def test_stream_statistics():
    z = P.randn(400_000, seed=7, draw=0)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs((z ** 3).mean()) < 2e-2 and abs((z ** 4).mean() - 3.0) < 5e-2
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 5e-3 and np.abs(z).max() < 6.8
    t, k = P.draw_rows(200_000, seed=7, draw=0, t_lo=100, t_hi=900, lambd=0.3)
    assert t.min() == 100 and t.max() == 899 and abs(t.mean() - 499.5) < 2.0
    assert abs(k.mean() - 0.7) < 4e-3                                                       # keep w.p. 1 - lambd
    _, k1 = P.draw_rows(1000, seed=7, draw=0, t_lo=0, t_hi=1, lambd=1.0)
    _, k0 = P.draw_rows(1000, seed=7, draw=0, t_lo=0, t_hi=1, lambd=0.0)
    assert not k1.any() and k0.all()                                                        # the reference's edge cases


def test_aux_uniform_stream():
    """EraseDiff's in-kernel uniform target: [0, 1), 24-bit, a function of the global index, independent of the
    noise domain of the same draw."""
    u = P.rand_aux(400_000, seed=7, draw=0)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 2e-3 and abs(u.var() - 1 / 12) < 1e-3
    assert np.array_equal(u * np.float32(2 ** 24), np.floor(u * np.float32(2 ** 24)))      # 24-bit grid
    assert np.array_equal(P.rand_aux(300, seed=7, draw=0, elem_offset=437), u[437:737])
    words, lane = P.noise_words(4000, seed=7, draw=0)
    noise_u = (np.take_along_axis(words, lane[:, None], axis=1)[:, 0] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    assert not np.array_equal(noise_u, u[:4000]) and abs(np.corrcoef(noise_u, u[:4000])[0, 1]) < 5e-2
