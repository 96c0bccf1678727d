"""numpy Philox4x32-10 (oracle; TEST INFRASTRUCTURE ONLY).

Restates the published algorithm (Salmon, Moraes, Dror, Shaw: "Parallel Random Numbers: As Easy as 1, 2, 3",
SC'11; Random123 ``philox.h``) and is pinned by the Random123 known-answer vectors (``kat_vectors``,
checked in tests/test_philox_ref.py).  It mirrors the counter/key convention of the product kernel
(``bez_isaacgym_b200/csrc/bezk_common.cuh: philox_reset_uniforms``): for env ``e`` at step ``s`` with seed ``k``
    ctr = (e_lo, e_hi, s_lo, (s_hi << 4) + j), key = (k_lo, k_hi), j = 0..8  ->  36 uniforms
    u = (x >> 8) * 2^-24 in [0, 1)
cols 0:18 feed the reset position draw and 18:36 the velocity draw (reference tasks/kick_env.py:786-787 draws
two (k,18) ``torch.rand`` blocks from the global generator; the product keys them by env id instead so
results do not depend on how many envs reset or on sharding -- documented deviation, SURVEY 5 "Seeding").
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def reset_uniforms(seed: int, step: int, n: int, first_env: int = 0) -> np.ndarray:
    """(n, 36) float32 uniforms for envs first_env .. first_env+n-1, identical to bezk_philox_uniforms."""
    env = np.arange(first_env, first_env + n, dtype=np.uint64)
    ctr = np.zeros((n, 9, 4), dtype=np.uint32)
    ctr[:, :, 0] = (env & MASK).astype(np.uint32)[:, None]
    ctr[:, :, 1] = (env >> np.uint64(32)).astype(np.uint32)[:, None]
    ctr[:, :, 2] = np.uint32(step & 0xFFFFFFFF)
    ctr[:, :, 3] = np.uint32(((step >> 32) << 4) & 0xFFFFFFFF) + np.arange(9, dtype=np.uint32)[None, :]
    key = np.zeros((n, 9, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    x = philox4x32_10(ctr, key).reshape(n, 36)
    return ((x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)
