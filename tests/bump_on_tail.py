"""BASELINE config 4: the bump-on-tail plasma of the reference's examples/bump-on-tail.toml (bulk + beam electrons against ions,
n_beam / n_0 = 0.03, v_beam = 0.25 c, sqrt(T_e / m_e) = 0.05 c, CFL 3, no filter), scaled to many particles on the bench grid.

Linear theory quoted in the example's header (examples/bump-on-tail.toml:4-15, non-relativistic electrostatic dispersion): the fastest
growing mode has Re(omega) = 1.0 omega_pe, Im(omega) = 0.075 omega_pe at k = 4.9 omega_pe / c, omega_pe the plasma frequency of the
bulk electrons.  `growth_rate_of_mode` fits the exponential phase of the spatial Fourier mode nearest that k, the way
examples/inference_two_stream.py:108-203 fits the field energy: least squares on the logarithm over a window of the linear phase.
"""
import numpy as np

C = 2.99792458e8
EPS0 = 8.85418782e-12
QE = 1.60217663e-19
ME = 9.10938371e-31
MP = 1.67262193e-27

VTH, GPDL, CFL = 0.07071067812, 2.565, 3.0     # examples/bump-on-tail.toml:21,45,51
V_BULK, V_BEAM = -2.25e6, 7.5e7                # :54, :117 (the bulk drifts back so that the net current vanishes)
BEAM_FRACTION = 0.03


def setup(n_total, G=70, length=None):
    """Species table and scalars for n_total macro-particles: half ions, half electrons split 97 : 3 into bulk and beam; one
    macro-particle weight for all, so density ratios are count ratios (the example scales grid_points_per_Debye_length instead).
    Default geometry = the example's own (1 m, 70 cells: the fastest-growing mode is mode 7, ten cells per wavelength).  A longer box
    at the same particles per cell is NOT equivalent: the reference's initial E_x = (dx / eps0) cumsum(rho_0) random-walks over the
    cells, so its noise grows like sqrt(G) and swamps the instability (saturation ~ 5e4 V/m) on a 4096-cell box below ~1e10 particles."""
    length = G / 70.0 if length is None else length   # the example's dx = 1 m / 70
    n_e = n_total // 2
    n_beam = int(round(n_e * BEAM_FRACTION / (1 + BEAM_FRACTION)))
    n_bulk, n_ion = n_e - n_beam, n_total - n_e
    w = EPS0 * ME * C ** 2 / QE ** 2 * G ** 2 / length / (2 * n_bulk) * VTH ** 2 * GPDL ** 2   # _state_initialization.py:172-185
    species = [dict(count=n_bulk, q=-QE * w, m=ME * w, qm=-QE / ME), dict(count=n_beam, q=-QE * w, m=ME * w, qm=-QE / ME),
               dict(count=n_ion, q=QE * w, m=MP * w, qm=QE / MP)]
    dx = length / G
    omega_pe = np.sqrt(n_bulk * w * QE ** 2 / (ME * EPS0 * length))   # _simulation.py:263-268 with the bulk electrons
    k = 4.9 * omega_pe / C
    return dict(G=G, length=length, dx=dx, dt=CFL * dx / C, species=species, omega_pe=omega_pe, k_theory=k,
                mode=int(round(k * length / (2 * np.pi))), gamma_theory=0.075 * omega_pe, counts=(n_bulk, n_beam, n_ion))


def particles_numpy(s, seed=250724):
    """(x0, v0, q, m, qm) in NumPy (for the oracle and for sharded runs: every rank can rebuild the same global arrays)."""
    rng = np.random.default_rng(seed)
    n_bulk, n_beam, n_ion = s["counts"]
    n = n_bulk + n_beam + n_ion
    L = s["length"]
    x0 = np.zeros((n, 3)); v0 = np.zeros((n, 3))
    x0[:, 0] = rng.uniform(-L / 2, L / 2, n)
    sig = VTH * C / np.sqrt(2)
    v0[:n_bulk, 0] = V_BULK + sig * rng.standard_normal(n_bulk)
    v0[n_bulk:n_bulk + n_beam, 0] = V_BEAM + sig * rng.standard_normal(n_beam)
    v0[n_bulk + n_beam:, 0] = sig * np.sqrt(ME / MP) * rng.standard_normal(n_ion)
    v0 = np.clip(v0, -0.99 * C, 0.99 * C)
    q = np.concatenate([np.full(sp["count"], sp["q"]) for sp in s["species"]])
    m = np.concatenate([np.full(sp["count"], sp["m"]) for sp in s["species"]])
    qm = np.concatenate([np.full(sp["count"], sp["qm"]) for sp in s["species"]])
    return x0, v0, q, m, qm


def particles_torch(s, device, seed=250724):
    """The same distributions drawn on the device (large runs)."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_bulk, n_beam, n_ion = s["counts"]
    n = n_bulk + n_beam + n_ion
    L = s["length"]
    x0 = torch.zeros((n, 3), dtype=torch.float64, device=device)
    v0 = torch.zeros((n, 3), dtype=torch.float64, device=device)
    x0[:, 0].uniform_(-L / 2, L / 2, generator=gen)
    sig = VTH * C / np.sqrt(2)
    v0[:, 0].normal_(0.0, 1.0, generator=gen)
    v0[:n_bulk, 0] *= sig; v0[:n_bulk, 0] += V_BULK
    v0[n_bulk:n_bulk + n_beam, 0] *= sig; v0[n_bulk:n_bulk + n_beam, 0] += V_BEAM
    v0[n_bulk + n_beam:, 0] *= sig * np.sqrt(ME / MP)
    v0.clamp_(-0.99 * C, 0.99 * C)
    return x0, v0


def growth_rate_of_mode(ex_hist, s, span=1, lo_factor=3.0, hi_factor=4.0):
    """ex_hist: (T, G) history of E_x.  Amplitude of the spatial Fourier modes within `span` of the theoretical one; the exponential
    phase is fitted by least squares on log amplitude between lo_factor x the initial (noise) level and 1 / hi_factor of the saturation
    amplitude -- the last time the amplitude is below the upper threshold before its maximum, and the last time it is below the lower
    one before that.  Returns (gamma [1/s], strongest mode number, (first, last) step of the fit window); gamma is NaN when the window
    holds fewer than 10 steps (too few particles: the noise floor is too close to the saturation level)."""
    ex = np.asarray(ex_hist, dtype=np.float64)
    T = ex.shape[0]
    spec = np.abs(np.fft.rfft(ex, axis=1))
    lo, hi = max(1, s["mode"] - span), min(spec.shape[1] - 1, s["mode"] + span)
    amp = np.sqrt((spec[:, lo:hi + 1] ** 2).sum(axis=1))        # the band around the fastest-growing mode
    best = lo + int(np.argmax(spec[:, lo:hi + 1].max(axis=0)))
    noise = np.median(amp[: max(8, T // 25)])
    t_top = int(np.argmax(amp))
    top = amp[t_top]
    under_hi = np.nonzero(amp[:t_top] <= top / hi_factor)[0]
    if len(under_hi) == 0:
        return float("nan"), best, (0, 0)
    b = int(under_hi[-1])
    under_lo = np.nonzero(amp[:b] <= lo_factor * noise)[0]
    a = int(under_lo[-1]) + 1 if len(under_lo) else 0
    if b - a < 10:
        return float("nan"), best, (a, b)
    t = np.arange(a, b + 1) * s["dt"]
    slope = np.polyfit(t, np.log(amp[a:b + 1]), 1)[0]
    return float(slope), best, (a, b)
