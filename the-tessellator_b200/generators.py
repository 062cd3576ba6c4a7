"""Seeded synthetic point sets for the parity tests and bench.py.

Counter-based RNG, all u64 wrapping arithmetic, so every language reproduces it bit for bit:

    mix(z): z += 0x9E3779B97F4A7C15; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EB; return z ^ (z >> 31)
    u(seed, i, c) = (mix(mix(seed) + 3*i + c) >> 11) * 2**-53            in [0, 1)

These are inputs only; they are handed unchanged to both the CUDA path and the CPU oracle.
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def mix(z: np.ndarray) -> np.ndarray:
    """splitmix64 output function on a uint64 array (wrapping)."""
    with np.errstate(over="ignore"):
        z = (z + _GOLD).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def u01(seed: int, counters: np.ndarray) -> np.ndarray:
    """u(seed, .) for explicit u64 counters (counter = 3*i + axis for coordinates)."""
    with np.errstate(over="ignore"):
        base = mix(np.array([seed], dtype=np.uint64))[0]
        bits = mix((base + counters.astype(np.uint64)).astype(np.uint64))
    return (bits >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def uniform(n: int, seed: int, start: int = 0) -> np.ndarray:
    """Points start..start+n-1 of the uniform [0,1)^3 stream `seed` -> (n, 3) float64."""
    i = np.arange(start, start + n, dtype=np.uint64)
    c = (np.uint64(3) * i)[:, None] + np.arange(3, dtype=np.uint64)[None, :]
    return u01(seed, c.reshape(-1)).reshape(n, 3)


def clustered(n: int, seed: int, k: int = 32, sigma: float = 0.02, background: float = 0.2) -> np.ndarray:
    """Config 4: `background` uniform + the rest in k isotropic Gaussians (sigma), centres uniform
    in [0.1, 0.9]^3, samples outside [0,1)^3 rejected and redrawn (Box-Muller on counters >= 3n)."""
    pts = uniform(n, seed)
    n_bg = int(round(background * n))
    # centres from counters 3n .. 3n+3k-1
    cen = u01(seed, np.arange(3 * n, 3 * n + 3 * k, dtype=np.uint64)).reshape(k, 3) * 0.8 + 0.1
    # cluster membership: point j (j >= n_bg) belongs to cluster (j - n_bg) % k
    todo = np.arange(n_bg, n, dtype=np.int64)
    ctr = np.uint64(3 * n + 3 * k)
    rounds = 0
    while todo.size:
        m = todo.size
        # 4 draws per pending point per round: two Box-Muller pairs -> 3 normals used
        base = ctr + np.uint64(4) * np.arange(m, dtype=np.uint64)
        u = u01(seed, (base[:, None] + np.arange(4, dtype=np.uint64)[None, :]).reshape(-1)).reshape(m, 4)
        ctr = ctr + np.uint64(4 * m)
        r1 = np.sqrt(-2.0 * np.log(1.0 - u[:, 0]))
        r2 = np.sqrt(-2.0 * np.log(1.0 - u[:, 2]))
        g = np.stack(
            [r1 * np.cos(2 * np.pi * u[:, 1]), r1 * np.sin(2 * np.pi * u[:, 1]), r2 * np.cos(2 * np.pi * u[:, 3])],
            axis=1,
        )
        cand = cen[(todo - n_bg) % k] + sigma * g
        ok = np.all((cand >= 0.0) & (cand < 1.0), axis=1)
        pts[todo[ok]] = cand[ok]
        todo = todo[~ok]
        rounds += 1
        if rounds > 64:
            raise RuntimeError("clustered(): rejection loop did not converge")
    return pts


def bcc(cells_per_side: int, seed: int, jitter: float = 1e-3, start: int = 0, count: int | None = None) -> np.ndarray:
    """Config 5: jittered BCC lattice, 2*m^3 points, a = 1/m; sites (i,j,k)a + a/4 and
    (i+.5,j+.5,k+.5)a + a/4; each coordinate += (2u-1)*jitter*a.  Point index p = 2*site + sub."""
    m = cells_per_side
    total = 2 * m ** 3
    if count is None:
        count = total - start
    p = np.arange(start, start + count, dtype=np.int64)
    site, sub = p // 2, p % 2
    i, j, k = site // (m * m), (site // m) % m, site % m
    a = 1.0 / m
    base = np.stack([i, j, k], axis=1).astype(np.float64) + 0.5 * sub[:, None].astype(np.float64)
    c = (np.uint64(3) * p.astype(np.uint64))[:, None] + np.arange(3, dtype=np.uint64)[None, :]
    u = u01(seed, c.reshape(-1)).reshape(count, 3)
    return base * a + a / 4 + (2.0 * u - 1.0) * (jitter * a)


def simple_cubic(m: int) -> np.ndarray:
    """Un-jittered simple-cubic lattice (exact-degeneracy robustness input, not part of the metric)."""
    g = (np.arange(m, dtype=np.float64) + 0.5) / m
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
