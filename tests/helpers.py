"""Small numpy helpers shared by the tests and the golden-fixture generator (no torch, no CUDA)."""
import numpy as np


def formula_table(n_entries: int, C: int, scale: float) -> np.ndarray:
    """Deterministic hash table: value(i) = ((i * 2654435761 mod 2^32) / 2^32 - 0.5) * 2 * scale, fp32.
    Big tables are never stored in fixtures; tests rebuild them from this closed form."""
    i = np.arange(n_entries * C, dtype=np.uint64)
    u = (i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    v = (u.astype(np.float64) / 4294967296.0 - 0.5) * 2.0 * scale
    return v.astype(np.float32).reshape(n_entries, C)


def make_rays(n, rng, near=0.90449, far=1.09551):
    """Cone-like rays through the +-0.15 m cube from a source 1 m away: [n, 8] = (o, d, near, far)."""
    ang = rng.uniform(0, 2 * np.pi, n)
    o = np.stack([np.cos(ang), np.sin(ang), rng.uniform(-0.02, 0.02, n)], -1)
    tgt = rng.uniform(-0.12, 0.12, (n, 3))
    d = tgt - o
    d = d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(1.0, 1.0002, (n, 1))
    return np.concatenate([o, d, np.full((n, 1), near), np.full((n, 1), far)], -1).astype(np.float32)
