"""Seeded synthetic plasmas shared by the parity tests (formulas of jaxincell/_state_initialization.py:51-85,172-185,259-264)."""
import numpy as np

from oracle import literal as L


def two_species(n_e, n_i, *, length, G, seed=1701, vth_e=0.05, drift=0.0, plus_minus=False, vth_yz=0.0,
                gpdl=2.0, amp=0.0, k=1.0, random_x=True, ion_mass=1.0, ion_vth_scale=1.0):
    """Electrons + ions with the reference's auto weight; returns dict of (N,3)/(N,) arrays and species table."""
    rng = np.random.default_rng(seed)
    c = L.speed_of_light

    def pos(n):
        x = rng.uniform(-length / 2, length / 2, n) if random_x else np.linspace(-length / 2, length / 2, n)
        x = x + amp * np.sin(k * 2 * np.pi / length * x)
        return np.stack([x, rng.uniform(-length / 2, length / 2, n), rng.uniform(-length / 2, length / 2, n)], axis=1)

    def vel(n, vth, vd, pm):
        v = np.stack([vth[a] * c / np.sqrt(2) * rng.standard_normal(n) for a in range(3)], axis=1)
        v[:, 0] += vd
        if pm:
            v[:, 0] *= (-1.0) ** np.arange(n)
        lim = 0.99 * c
        return np.where(np.abs(v) >= lim, np.sign(v) * lim, v)

    xe, xi = pos(n_e), pos(n_i)
    ve = vel(n_e, (vth_e, vth_yz, vth_yz), drift, plus_minus)
    mi = ion_mass * L.mass_proton
    vthi = ion_vth_scale * vth_e * np.sqrt(L.mass_electron / mi)
    vi = vel(n_i, (vthi, vthi if vth_yz else 0.0, vthi if vth_yz else 0.0), 0.0, False)
    qe, qi = -L.elementary_charge, L.elementary_charge

    def weight(n):  # _state_initialization.py:172-185
        return L.epsilon_0 * L.mass_electron * c ** 2 / qe ** 2 * G ** 2 / length / (2 * n) * max(vth_e, vth_yz) ** 2 * gpdl ** 2

    we, wi = weight(n_e), weight(n_i)
    q = np.concatenate([np.full(n_e, qe * we), np.full(n_i, qi * wi)])
    m = np.concatenate([np.full(n_e, L.mass_electron * we), np.full(n_i, mi * wi)])
    qm = np.concatenate([np.full(n_e, qe / L.mass_electron), np.full(n_i, qi / mi)])
    species = [dict(count=n_e, q=qe * we, m=L.mass_electron * we, qm=qe / L.mass_electron),
               dict(count=n_i, q=qi * wi, m=mi * wi, qm=qi / mi)]
    return dict(x0=np.concatenate([xe, xi]), v0=np.concatenate([ve, vi]), q=q, m=m, qm=qm, species=species)


def cfl_dt(length, G, cfl):
    return cfl * (length / G) / L.speed_of_light
