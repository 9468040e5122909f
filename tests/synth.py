"""Seeded synthetic inputs shared by the golden generator, the CPU tests, the GPU parity tests and bench.py.

Everything comes from numpy's PCG64 `default_rng(seed)` so the same arrays can be regenerated anywhere without the
reference (random-init weights with torch's default init *distributions*, N(0,1) tokens,
Bernoulli(0.5) missing-modality masks with token 0 always present).
"""
import numpy as np

F32 = np.float32


def decoder_inputs(N: int, D: int, L: int, seed: int = 0, symmetric: bool = True, unit_scale: bool = True):
    """z [N,D] ~ N(0,1)/sqrt(D) (SURVEY §8d decoder-only timing), P [L,D,D] ~ U(+-1/sqrt(D)) (nn.Bilinear init)."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((N, D), dtype=F32)
    if unit_scale:
        z /= F32(np.sqrt(D))
    bound = F32(1.0 / np.sqrt(D))
    P = rng.uniform(-bound, bound, size=(L, D, D)).astype(F32)
    if symmetric:
        P = np.triu(P) + np.swapaxes(np.triu(P, 1), -1, -2)
    return z, P
