"""Seeded synthetic datasets for the search hot path.

The reference ships no data; its README example draws IID Gaussians (README.md:95-97,135), on which
recall@10 >= 0.95 is unreachable at sane ef.  The headline generator is therefore
"latent": a rank-r Gaussian latent model plus isotropic noise, which behaves like real embeddings.

    G0  "iid"         x ~ N(0, I_D)
    G1  "latent"      x = z A + sigma * eps,  A in R^{r x D}, A_ij ~ N(0, 1/r), z ~ N(0, I_r)
    G1n "latent-norm" G1, rows L2-normalised            (inner-product / "angular" configs)
    G1u "latent-u8"   G1, affine map of [-4 s, 4 s] to [0, 255], rounded, clipped   (uint8 configs)
    G1i "latent-i8"   same, mapped to [-128, 127]

Seeds: mixing matrix 41, database 42, queries 43 (different streams of the same law).
"""
from __future__ import annotations

import numpy as np

GENERATORS = ("iid", "latent", "latent-norm", "latent-u8", "latent-i8")
A_SEED, DATA_SEED, QUERY_SEED = 41, 42, 43


def _mixing(dim: int, rank: int) -> np.ndarray:
    rng = np.random.default_rng(A_SEED)
    return (rng.standard_normal((rank, dim)) / np.sqrt(rank)).astype(np.float32)


def _draw(gen: str, n: int, dim: int, seed: int, rank: int, sigma: float, chunk: int = 1 << 18) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), dtype=np.float32)
    A = None if gen == "iid" else _mixing(dim, rank)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        if gen == "iid":
            out[lo:hi] = rng.standard_normal((hi - lo, dim), dtype=np.float32)
        else:
            z = rng.standard_normal((hi - lo, rank), dtype=np.float32)
            eps = rng.standard_normal((hi - lo, dim), dtype=np.float32)
            out[lo:hi] = z @ A + np.float32(sigma) * eps
    return out


def make(gen: str, n: int, dim: int, *, queries: bool = False, rank: int = 16, sigma: float = 0.1,
         stream: int = 0) -> np.ndarray:
    """Return `n` vectors of generator `gen` (database stream, or the query stream if `queries`).
    `stream` > 0 selects an independent draw of the same law (dataset shards)."""
    if gen not in GENERATORS:
        raise ValueError(f"unknown generator {gen!r}; expected one of {GENERATORS}")
    seed = (QUERY_SEED if queries else DATA_SEED) + 7919 * stream
    base = "iid" if gen == "iid" else "latent"
    x = _draw(base, n, dim, seed, rank, sigma)
    if gen == "latent-norm":
        x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-30).astype(np.float32)
    elif gen in ("latent-u8", "latent-i8"):
        # per-coordinate std of G1 is sqrt(||A_col||^2 + sigma^2) ~ sqrt(1 + sigma^2); use one global scale
        s = float(np.sqrt(1.0 + sigma * sigma))
        y = (x + 4.0 * s) * (255.0 / (8.0 * s))
        y = np.clip(np.rint(y), 0, 255)
        if gen == "latent-u8":
            return y.astype(np.uint8)
        return (y - 128.0).astype(np.int8)
    return x


def make_device(gen: str, n: int, dim: int, *, queries: bool = False, rank: int = 16, sigma: float = 0.1,
                stream: int = 0, device=None, chunk: int = 1 << 20):
    """The same laws drawn on the GPU with torch's generator (a torch tensor on `device`): for the large benchmark
    legs, where the numpy draw of 10M+ rows takes longer than building the graph.  Same mixing matrix as `make`;
    the random streams differ from numpy's, so `make` and `make_device` give different (equally distributed) rows."""
    import torch
    if gen not in GENERATORS:
        raise ValueError(f"unknown generator {gen!r}; expected one of {GENERATORS}")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    g = torch.Generator(device=device)
    g.manual_seed((QUERY_SEED if queries else DATA_SEED) + 7919 * stream)
    out_dtype = {"latent-u8": torch.uint8, "latent-i8": torch.int8}.get(gen, torch.float32)
    out = torch.empty((n, dim), dtype=out_dtype, device=device)
    A = None if gen == "iid" else torch.from_numpy(_mixing(dim, rank)).to(device)
    s = float(np.sqrt(1.0 + sigma * sigma))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        if gen == "iid":
            x = torch.randn((hi - lo, dim), generator=g, device=device, dtype=torch.float32)
        else:
            z = torch.randn((hi - lo, rank), generator=g, device=device, dtype=torch.float32)
            x = z @ A + sigma * torch.randn((hi - lo, dim), generator=g, device=device, dtype=torch.float32)
        if gen == "latent-norm":
            x = x / torch.clamp(torch.linalg.norm(x, dim=1, keepdim=True), min=1e-30)
        elif gen in ("latent-u8", "latent-i8"):
            x = torch.clamp(torch.round((x + 4.0 * s) * (255.0 / (8.0 * s))), 0, 255)
            if gen == "latent-i8":
                x = x - 128.0
        out[lo:hi] = x.to(out_dtype)
    return out


def dtype_code(arr_or_dtype) -> str:
    dt = np.dtype(getattr(arr_or_dtype, "dtype", arr_or_dtype))
    return {np.dtype(np.float32): "f32", np.dtype(np.uint8): "u8", np.dtype(np.int8): "i8"}[dt]
