"""Seeded inputs of the isolated-particle step tests (shared by the golden generator, the oracle pin and
the GPU parity test): a disk with a tail of eccentric / hyperbolic / circular orbits."""
import numpy as np

from gplum_b200 import disk


def make_case(n=4000, seed=3, t0=0.25):
    rng = np.random.default_rng(seed)
    d = disk.make_disk(n, a_in=0.6, a_out=2.5, seed=seed)
    pos, vel = d["pos"].copy(), d["vel"].copy()
    # stir: moderate and high eccentricities (some beyond the 0.8 switch), a few exactly circular, a few unbound
    k = n // 8
    vel[:k] *= (0.55 + 0.9 * rng.random(k))[:, None]
    r = np.sqrt((pos[k:k + 20] ** 2).sum(1))
    vel[k:k + 20] = np.stack([-pos[k:k + 20, 1], pos[k:k + 20, 0], np.zeros(20)], 1) / r[:, None] ** 1.5
    pos[k:k + 20, 2] = 0.0
    vel[k + 20:k + 30] *= 1.6
    acc = rng.normal(size=(n, 3)) * 1e-4
    acc0 = np.abs(rng.normal(size=n)) * 1e-4
    isolated = (rng.random(n) < 0.9).astype(np.int32)
    time = np.full(n, t0)
    dt = np.where(rng.random(n) < 0.5, 0.0, 2.0 ** -rng.integers(7, 14, n))
    return {"pos": pos, "vel": vel, "acc": acc, "acc0": acc0, "isolated": isolated, "time": time, "dt": dt,
            "t0": t0, "t1": t0 + 2.0 ** -6}
