"""Deterministic synthetic PET sets of the shapes BASELINE.json names.

Per chromosome of length L: a background (1/d distance law on [1e2, 2e6] plus a self-ligation lump,
lognormal with median ~400 bp, 25 % of background PETs) and loops (PETs scattered N(0, sigma) around
anchor pairs with spans U[2e4, 2e6]).  Rows are shuffled; id = row.  X <= Y as the reference's PET
class guarantees (cLoops/io.py:51-54).
"""
from __future__ import annotations

import numpy as np

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
        133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
        58617616, 64444167, 46709983, 50818468, 156040895]
CHROMS = ["chr%d" % i for i in range(1, 23)] + ["chrX"]


def chromosome(n: int, length: int, seed: int, loop_frac: float = 0.08, sigma: float = 500.0, pets_per_loop: int = 40):
    """Returns int32 X, Y (row order)."""
    rng = np.random.default_rng(seed)
    n_loop = int(n * loop_frac)
    n_bg = n - n_loop
    n_self = n_bg // 4
    n_far = n_bg - n_self
    x_bg = rng.integers(0, length, n_bg)
    d_far = np.exp(rng.uniform(np.log(1e2), np.log(2e6), n_far))
    d_self = rng.lognormal(np.log(400.0), 0.6, n_self)
    d = np.concatenate([d_far, d_self]).astype(np.int64)
    y_bg = np.minimum(x_bg + d, length - 1)
    k = max(1, n_loop // pets_per_loop)
    span = rng.integers(20000, 2000000, k)
    ax = rng.integers(0, max(1, length - 2000001), k)
    ay = ax + span
    which = rng.integers(0, k, n_loop)
    x_lp = (ax[which] + rng.normal(0, sigma, n_loop)).astype(np.int64)
    y_lp = (ay[which] + rng.normal(0, sigma, n_loop)).astype(np.int64)
    X = np.concatenate([x_bg, x_lp])
    Y = np.concatenate([y_bg, y_lp])
    X = np.clip(X, 0, length - 1)
    Y = np.clip(Y, 0, length - 1)
    X, Y = np.minimum(X, Y), np.maximum(X, Y)
    p = rng.permutation(n)
    return X[p].astype(np.int32), Y[p].astype(np.int32)


def config2(n: int = 10_000_000, seed: int = 20240 + 200):
    """'synthetic ChIA-PET 10M cis PETs, single chromosome, eps=1000 minPts=5' (BASELINE.json configs[1])."""
    return chromosome(n, 249_000_000, seed, loop_frac=0.08, sigma=500.0)


GENOME_PARAMS = {3: dict(sigma=1500.0, loop_frac=0.06), 4: dict(sigma=1500.0, loop_frac=0.06)}


def genome_counts(n_total: int):
    """PETs per chromosome, proportional to the hg38 lengths (configs 3 and 4)."""
    tot = float(sum(HG38))
    return [int(round(n_total * L / tot)) for L in HG38]


def genome_chrom(n_total: int, config: int, ci: int, sigma: float = 1500.0, loop_frac: float = 0.06):
    """Chromosome ``ci`` (0-based, chr1..22, X) of the ``n_total``-PET genome of config 3 / 4: (name, X, Y)."""
    n = genome_counts(n_total)[ci]
    return (CHROMS[ci], *chromosome(n, HG38[ci], 20240 + config * 100 + ci, loop_frac=loop_frac, sigma=sigma))


def genome(n_total: int, config: int, sigma: float = 1500.0, loop_frac: float = 0.06, only=None):
    """Per-chromosome sets with PET counts proportional to hg38 lengths (configs 3 and 4); ``only``: the chromosome
    indices wanted (a rank's shard)."""
    return [genome_chrom(n_total, config, ci, sigma, loop_frac) for ci in range(len(HG38)) if only is None or ci in only]

HIC_PETS_PER_KB = 66.0     # config 4: chr1 holds 16.4 M of the 200 M PETs on 249 Mb


def hic_density(n: int, seed: int, sigma: float = 1500.0, loop_frac: float = 0.06):
    """A chromosome of ``n`` PETs at the PET density of config 4 (deep Hi-C): length = n / 66 PETs per kb."""
    length = int(n * 1000.0 / HIC_PETS_PER_KB)
    return chromosome(n, length, seed, loop_frac=loop_frac, sigma=sigma)
