"""Synthetic inputs of the BASELINE.json configs: U[0,1]^2 unique locations per row, fixed outcome
proportions, p = 3 covariates, 10 % missing, numpy.random.default_rng(2021)."""
import numpy as np

CONFIGS = {
    # name: (q, n, proportions)
    "C1": (1, 625, (1.0,)),
    "C2": (1, 10_000, (1.0,)),
    "C3": (3, 300_000, (1 / 3, 1 / 3, 1 / 3)),
    "C4": (3, 1_000_000, (.6, .3, .1)),
    "C5": (5, 4_000_000, (.55, .25, .10, .07, .03)),
}
C3_OBSERVED_FRACTION = (0.9, 0.5, 0.2)

# fixed-theta parity points
THETA_Q1 = np.array([2.3, 1.0, 1.0, 6.0])
THETA_Q2 = np.array([1.0, 1.5, .1, .51, 1.0, 2.0, 5.0, 1.0])
THETA_Q3 = np.array([1.0, 1.5, .8, .1, .51, .3, 1.0, 2.0, 3.0, 2.0, .5, 5.0, 1.0, 2.0, 1.5])
THETA_Q5 = np.array([1.0, 1.5, .8, 1.2, .9, .1, .51, .3, .2, .4, 1.0, 2.0, 3.0, 1.5, 2.5, 2.0, .5, 5.0,
                     1.0, 2.0, 1.5, 1.2, .8, 1.7, 2.2, .6, 1.1, .9])


def theta_for(q):
    return {1: THETA_Q1, 2: THETA_Q2, 3: THETA_Q3, 5: THETA_Q5}[q].copy()


def make_data(q, n, proportions=None, missing=0.1, seed=2021, observed_fraction=None):
    """Returns rows sorted by (Var1, Var2) as spamtree() passes them on (R/spamtree_fit.R:214)."""
    rng = np.random.default_rng(seed)
    proportions = (1.0 / q,) * q if proportions is None else proportions
    coords = rng.random((n, 2))
    mv_id = rng.choice(np.arange(1, q + 1), size=n, p=np.asarray(proportions) / np.sum(proportions)).astype(np.int64)
    if q > 1:
        for j in range(1, q + 1):  # every outcome present
            mv_id[j - 1] = j
    X = rng.standard_normal((n, 3))
    beta = np.array([-1.0, .5, 1.0])
    field = np.sin(4 * coords[:, 0]) * np.cos(3 * coords[:, 1]) + 0.3 * mv_id
    y = X @ beta + field + np.sqrt(.1) * rng.standard_normal(n)
    if observed_fraction is not None:
        miss = rng.random(n) > np.asarray(observed_fraction)[mv_id - 1]
    else:
        miss = rng.random(n) < missing
    y = y.copy()
    y[miss] = np.nan
    order = np.lexsort((np.arange(n), coords[:, 1], coords[:, 0]))
    return {"y": y[order], "X": X[order], "coords": coords[order], "mv_id": mv_id[order], "q": q, "n": n}


def make_config(name, n=None, seed=2021):
    q, n0, prop = CONFIGS[name]
    return make_data(q, n or n0, prop, seed=seed, observed_fraction=C3_OBSERVED_FRACTION if name == "C3" else None)


def default_bounds(q, btmlim=1e-3, toplim=1e3):
    """theta bounds of spamtree() (R/spamtree_fit.R:106-134)"""
    n_cbase = 3 if q > 2 else 1
    npars = 3 * q + n_cbase
    b = np.zeros((npars, 2))
    b[:, 0], b[:, 1] = btmlim, toplim
    if q > 1:
        b[1:q, 0] = -toplim
    if n_cbase == 3:
        b[npars - 2, :] = (btmlim, 1 - btmlim)
    k = q * (q - 1) // 2
    if q > 1:
        vb = np.zeros((k, 2))
        vb[:, 0], vb[:, 1] = btmlim, toplim - btmlim
        b = np.vstack([b, vb])
    return b
