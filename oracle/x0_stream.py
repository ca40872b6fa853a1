"""TEST INFRASTRUCTURE — bit-exact NumPy twin of the device-side counter-based state generator
(q_learning_with_hjb_b200/csrc/sample.cu, ``hjb_sample_states``): Philox4x32-10 keyed by the seed and counted by the
global sample index, x = wrap(fma(std, t, mean)) with t = 2 (r >> 8) 2^-24 - 1.

The distribution is the reference's x0 = wrap(U(-x0_std, x0_std) + x0_mean) (dynamics/dynamics_basic.py:28-29); the
STREAM is new (the reference's global NumPy RNG is serial), so this twin is what lets the oracle consume exactly the
numbers the device generated.  Only tests/, bench.py's checker legs and __graft_entry__.smoke() import it."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
ANGLES = {"linear": (), "cartpole": (1,), "acrobot": (0, 1), "quad2d": (2,), "quad10d": (3, 4)}


def philox4x32_10(k0: int, k1: int, c0, c1, c2, c3):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11): uint32 arrays in, four uint32 arrays out."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _fma32(a, b, c):
    """fmaf on float32 arrays: products of two float32 are exact in float64, one rounding to float32 at the end."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def wrap_pi_f32(a):
    """hjb_common.cuh::wrap_pi in emulated float32 arithmetic."""
    inv2pi, hi, lo = np.float32(0.15915494309189533577), np.float32(6.28318548202514648438), np.float32(-1.74845553146951715e-07)
    k = np.floor(_fma32(a, inv2pi, np.float32(0.5)))
    return _fma32(-k, lo, _fma32(-k, hi, a))


def sample_states(sys_kind: str, mean, std, seed: int, first: int, count: int) -> np.ndarray:
    """float32 [count, n]: samples first .. first + count - 1 of the stream keyed by ``seed``."""
    mean, std = np.asarray(mean, dtype=np.float32), np.asarray(std, dtype=np.float32)
    n = mean.shape[0]
    g = np.arange(first, first + count, dtype=np.uint64)
    x = np.empty((count, n), dtype=np.float32)
    for q in range((n + 3) // 4):
        r = philox4x32_10(seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, g & MASK, g >> np.uint64(32), q, 0)
        for j in range(4):
            c = 4 * q + j
            if c < n:
                t = _fma32((r[j] >> np.uint32(8)).astype(np.float32), np.float32(2.0 ** -23), np.float32(-1.0))
                x[:, c] = _fma32(std[c], t, mean[c])
    for c in ANGLES[sys_kind]:
        x[:, c] = wrap_pi_f32(x[:, c])
    return x
