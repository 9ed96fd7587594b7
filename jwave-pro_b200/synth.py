"""Synthetic input generators shared by tests and bench: counter-based uniform(-1,1) and a chirp family."""
import numpy as np


def splitmix_uniform(seed, shape):
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        z = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + np.uint64(seed)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)
    return (u * 2.0 - 1.0).reshape(shape)


def chirp(batch, n):
    t = np.arange(n, dtype=np.float64) / n
    f0 = 4.0 + (np.arange(batch) % 13)[:, None]
    k = n / 8.0
    return np.sin(2.0 * np.pi * (f0 * t[None, :] + 0.5 * k * t[None, :] ** 2))
