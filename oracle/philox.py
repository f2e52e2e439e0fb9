"""TEST INFRASTRUCTURE (oracle): numpy restatement of the counter-based random streams the CUDA chain kernels draw
from (quinn_b200/csrc/qb_device.cuh: qb_philox / qb_rand4 / qb_u01 / qb_normal4), so that a Philox-driven device chain
can be replayed step by step on the CPU.  The reference itself draws from numpy's global Mersenne Twister
(admcmc.py:70, hmc.py:43, mcmc.py:75); these streams replace it in the many-chain mode, keyed so that results do not
depend on how chains are sharded.  Only tests/ may import this.

Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi ^ 0x5851F42D),
counter = (index/4, step_lo, chain_lo, stream | step_hi<<8 | chain_hi<<16); Box-Muller on pairs of 32-bit words.
"""
import numpy as np

STREAM_INCR, STREAM_Z0, STREAM_UNIF, STREAM_VI = 0, 1, 2, 3
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: 4 arrays of uint32 (broadcastable), key: 2 ints.  Returns 4 uint32 arrays."""
    c = [np.asarray(v, dtype=np.uint64) & _MASK for v in ctr]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c = [(hi1 ^ c[1] ^ np.uint64(k0)) & _MASK, lo1, (hi0 ^ c[3] ^ np.uint64(k1)) & _MASK, lo0]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [v.astype(np.uint32) for v in c]


def rand4(seed, chain, step, stream, idx4):
    """qb_rand4: four 32-bit words for (seed, global chain id, step, stream, index/4)."""
    seed, chain, step = int(seed) & (2 ** 64 - 1), int(chain), int(step)
    key = (seed & 0xFFFFFFFF, ((seed >> 32) & 0xFFFFFFFF) ^ 0x5851F42D)
    w = (stream & 0xFF) | (((step >> 32) & 0xFF) << 8) | (((chain >> 32) & 0xFFFF) << 16)
    idx4 = np.asarray(idx4, dtype=np.uint64)
    return philox4x32_10((idx4, np.uint64(step & 0xFFFFFFFF), np.uint64(chain & 0xFFFFFFFF), np.uint64(w)), key)


def u01(a):
    """qb_u01: (a + 0.5) * 2^-32 in double."""
    return (np.asarray(a, dtype=np.float64) + 0.5) * 2.3283064365386963e-10


def normals(seed, chain, step, stream, n):
    """The first n standard normals of a stream in double precision (qb_normal4, double overload): element i comes
    from block i // 4, position i % 4 = (r1 cos, r1 sin, r2 cos, r2 sin)."""
    nb = (n + 3) // 4
    r = rand4(seed, chain, step, stream, np.arange(nb))
    r1 = np.sqrt(-2.0 * np.log(u01(r[0])))
    r2 = np.sqrt(-2.0 * np.log(u01(r[2])))
    a1, a2 = 2.0 * np.pi * u01(r[1]), 2.0 * np.pi * u01(r[3])
    z = np.stack([r1 * np.cos(a1), r1 * np.sin(a1), r2 * np.cos(a2), r2 * np.sin(a2)], axis=1).reshape(-1)
    return z[:n]


def uniform(seed, chain, step):
    """The accept/reject uniform of a step (qb_mh_step)."""
    return float(u01(rand4(seed, chain, step, STREAM_UNIF, np.zeros(1))[0])[0])


def amcmc_draws(seed, chain, nsteps, P, t_start=0):
    """Draws an AMCMC chain consumes in Philox mode: z0[t] (common-mode normal of the rank-1 initial covariance),
    z[t, P] and u[t]."""
    z0 = np.array([normals(seed, chain, t_start + t, STREAM_Z0, 1)[0] for t in range(nsteps)])
    z = np.stack([normals(seed, chain, t_start + t, STREAM_INCR, P) for t in range(nsteps)])
    u = np.array([uniform(seed, chain, t_start + t) for t in range(nsteps)])
    return dict(z0=z0, z=z, u=u)


def hmc_draws(seed, chain, nsteps, P, t_start=0):
    """Momentum draws p[t, P] and uniforms u[t] of an HMC / MALA chain in Philox mode."""
    p = np.stack([normals(seed, chain, t_start + t, STREAM_INCR, P) for t in range(nsteps)])
    u = np.array([uniform(seed, chain, t_start + t) for t in range(nsteps)])
    return dict(p=p, u=u)
