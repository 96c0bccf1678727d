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


def normals18(seed: int, step: int, envs) -> np.ndarray:
    """(len(envs), 18) float32 standard normals identical (to libm rounding) to ``bezk_normal_noise`` /
    the Philox path of ``bezk_policy_head`` (``bez_isaacgym_b200/csrc/bezk_rollout.cu: philox_normals18``):
    key = (seed_lo ^ 0x5851F42D, seed_hi), ctr = (e_lo, e_hi, s_lo, (s_hi << 4) + j), j = 0..4; each block's
    (x, y) and (z, w) feed one Box-Muller pair: u1 = ((a >> 8) + 1) * 2^-24 in (0, 1], u2 = (b >> 8) * 2^-24,
    r = sqrt(-2 ln u1), (r cos 2 pi u2, r sin 2 pi u2).  Replaces torch's Normal.sample() draw (documented
    deviation: keyed by env id, so sharding and batch size do not change the noise)."""
    env = np.asarray(list(envs), dtype=np.uint64)
    n = env.shape[0]
    ctr = np.zeros((n, 5, 4), dtype=np.uint32)
    ctr[:, :, 0] = (env & MASK).astype(np.uint32)[:, None]
    ctr[:, :, 1] = (env >> np.uint64(32)).astype(np.uint32)[:, None]
    ctr[:, :, 2] = np.uint32(step & 0xFFFFFFFF)
    ctr[:, :, 3] = np.uint32(((step >> 32) << 4) & 0xFFFFFFFF) + np.arange(5, dtype=np.uint32)[None, :]
    key = np.zeros((n, 5, 2), dtype=np.uint32)
    key[..., 0] = np.uint32((seed & 0xFFFFFFFF) ^ 0x5851F42D)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    x = philox4x32_10(ctr, key).reshape(n, 10, 2)           # 10 (a, b) pairs per env
    scale = np.float32(1.0 / 16777216.0)
    u1 = ((x[..., 0] >> np.uint32(8)) + np.uint32(1)).astype(np.float32) * scale
    u2 = (x[..., 1] >> np.uint32(8)).astype(np.float32) * scale
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    th = (np.float32(6.283185307179586) * u2).astype(np.float32)
    z = np.stack((r * np.cos(th), r * np.sin(th)), axis=-1).astype(np.float32).reshape(n, 20)
    return z[:, :18]
