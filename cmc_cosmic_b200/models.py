"""Synthetic initial conditions in Henon units (G = M = 1, E = -1/4): Plummer sphere and King model.

These stand in for the reference's generators (COSMIC sample/cmc/elson.py, king.py), which cannot travel to the GPU
box; they produce the same families (r, v_r, v_t) but not the same random realisation.  Parity fixtures made WITH the
reference's generators are committed under tests/golden/.
"""
import numpy as np


def _isotropic_split(v, rng):
    """|v| -> (v_r, v_t) for an isotropic direction."""
    cos_t = rng.uniform(-1.0, 1.0, v.shape)
    return v * cos_t, v * np.sqrt(1.0 - cos_t * cos_t)


def plummer(N, seed=1234, r_max=100.0):
    """Plummer sphere (Aarseth, Henon & Wielen 1974 sampling), truncated at r_max virial radii; equal masses 1/N."""
    rng = np.random.default_rng(seed)
    a = 3.0 * np.pi / 16.0                      # Plummer scale length in Henon units
    m_max = (r_max / a) ** 3 / (1.0 + (r_max / a) ** 2) ** 1.5
    u = rng.uniform(0.0, m_max, N)
    u = np.clip(u, 1e-300, None)
    x = 1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0)  # radius in units of a
    # speed: q = v / v_esc from g(q) = q^2 (1 - q^2)^(7/2) by rejection
    q = np.empty(N)
    todo = np.arange(N)
    while todo.size:
        q1 = rng.uniform(0.0, 1.0, todo.size)
        q2 = rng.uniform(0.0, 0.1, todo.size)
        ok = q2 < q1 * q1 * (1.0 - q1 * q1) ** 3.5
        q[todo[ok]] = q1[ok]
        todo = todo[~ok]
    v = q * np.sqrt(2.0) * (1.0 + x * x) ** (-0.25)
    r = x * a
    v = v / np.sqrt(a)
    vr, vt = _isotropic_split(v, rng)
    order = np.argsort(r, kind="stable")
    return r[order], vr[order], vt[order], np.full(N, 1.0 / N)


def _king_profile(w0, n_steps=20000):
    """Integrate the King (1966) model: returns x (in core radii), W(x), enclosed mass (arbitrary units)."""
    from math import erf, exp, pi, sqrt

    def rho(w):
        if w <= 0.0:
            return 0.0
        return exp(w) * erf(sqrt(w)) - sqrt(4.0 * w / pi) * (1.0 + 2.0 * w / 3.0)

    rho0 = rho(w0)
    xs, ws, ms = [0.0], [w0], [0.0]
    x, w, dw, m = 1e-6, w0 - 1.5 * 1e-12, -3.0 * 1e-6, 0.0
    h = 1e-3

    def f(x, w, dw):
        return dw, -9.0 * rho(w) / rho0 - 2.0 * dw / x

    while w > 0.0 and len(xs) < 4000000:
        k1 = f(x, w, dw)
        k2 = f(x + h / 2, w + h / 2 * k1[0], dw + h / 2 * k1[1])
        k3 = f(x + h / 2, w + h / 2 * k2[0], dw + h / 2 * k2[1])
        k4 = f(x + h, w + h * k3[0], dw + h * k3[1])
        w_new = w + h / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
        dw_new = dw + h / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
        if w_new <= 0.0:
            frac = w / (w - w_new)
            x += frac * h
            xs.append(x); ws.append(0.0); ms.append(-x * x * (dw + frac * (dw_new - dw)))
            break
        x += h
        w, dw = w_new, dw_new
        xs.append(x); ws.append(w); ms.append(-x * x * dw)   # M(x) ~ -x^2 dW/dx
        h = min(max(1e-3, 2e-3 * x), 0.05)
    return np.array(xs), np.array(ws), np.array(ms)


def king(N, w0=5.0, seed=12345):
    """King model of dimensionless central potential w0; equal masses 1/N; scaled to Henon units."""
    rng = np.random.default_rng(seed)
    x, w, m = _king_profile(w0)
    m = m / m[-1]
    m = np.maximum.accumulate(m)
    u = rng.uniform(0.0, 1.0, N)
    r = np.interp(u, m, x)
    wr = np.interp(r, x, w)
    # speeds: f(v) ~ v^2 (exp(W - v^2/2) - 1), v in units of sigma, 0 <= v < sqrt(2W)
    v = np.empty(N)
    todo = np.arange(N)
    while todo.size:
        ww = wr[todo]
        vmax = np.sqrt(2.0 * ww)
        v1 = rng.uniform(0.0, 1.0, todo.size) * vmax
        fmax = vmax * vmax * np.expm1(ww) + 1e-300     # crude but valid envelope: v^2 <= vmax^2, (e^(W-v^2/2) - 1) <= e^W - 1
        f1 = v1 * v1 * np.expm1(np.maximum(ww - 0.5 * v1 * v1, 0.0))
        ok = rng.uniform(0.0, 1.0, todo.size) * fmax < f1
        v[todo[ok]] = v1[ok]
        todo = todo[~ok]
    # model units: G = 1, core radius r_c = 1, 4 pi G rho0 r_c^2 = 9 sigma^2.  Rescale to G = M = 1, E = -1/4 via the
    # sampled system's own energies (virial ratio is preserved by the King DF).
    order = np.argsort(r, kind="stable")
    r, v = r[order], v[order]
    mass = np.full(N, 1.0 / N)
    # potential energy of equal-mass shells: W_pot = -sum_i m_i * (M(<r_i)/r_i + sum_{j>i} m_j / r_j)
    cum_in = (np.arange(N) + 0.5) / N
    outer = np.concatenate([np.cumsum((mass / r)[::-1])[::-1][1:], [0.0]])
    pot = -np.sum(mass * (cum_in / r + outer)) * 0.5
    kin_sigma2 = 0.5 * np.sum(mass * v * v)             # in units of sigma^2
    # choose sigma^2 so that 2K + W = 0, then scale lengths so that E = K + W = -1/4
    sigma2 = -pot / (2.0 * kin_sigma2)
    e_tot = kin_sigma2 * sigma2 + pot                     # = pot / 2
    lam = e_tot / (-0.25)                                 # r -> r * lam  (W scales as 1/lam)
    r = r * lam
    v = v * np.sqrt(sigma2 / lam)
    vr, vt = _isotropic_split(v, rng)
    return r, vr, vt, mass


def kroupa_like_masses(N, seed=7, m_lo=0.08, m_hi=150.0, alpha=2.3):
    """Power-law masses normalised to sum 1 (stand-in for the reference's IMF sampler; mass spectrum only)."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(0.0, 1.0, N)
    a1 = 1.0 - alpha
    m = (m_lo ** a1 + u * (m_hi ** a1 - m_lo ** a1)) ** (1.0 / a1)
    return m / m.sum()
