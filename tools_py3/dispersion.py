"""Kinetic dispersion relation of a 1-D electrostatic Vlasov-Poisson plasma made of (shifted) Maxwellian species --
py3 restatement of what the reference's analysis tool solves (/root/reference/tools/dispersion.py:130-157: the
dispersion function; :33-60: Muller's method; :67-69: the default starting guesses).  It supplies the analytic numbers
the physics acceptance tests compare against: the bump-on-tail growth rate at k = 0.36 (the default input,
src/pic1dp_input.F90:47-72) and the Landau damping rate of a thermal plasma at k = 0.5 (configs[2]).

    D(omega) = 1 + sum_s  n_s Z_s^2 / m_s / (k^2 vth_s^2) * (1 + zeta_s Z(zeta_s)),
    zeta_s = (omega / k - v0_s) / sqrt(2 vth_s^2),  vth_s^2 = T_s / m_s,  Z = plasma dispersion function.

Test / analysis infrastructure, not part of the product.

    python -m tools_py3.dispersion            # prints the two roots used by the tests
"""
from __future__ import annotations

import cmath
import math
from typing import Callable, Sequence, Tuple

from scipy.special import wofz

Species = Tuple[float, float, float, float, float]  # charge Z, mass m, temperature T, density n, drift v0

# starting points of the root search (the reference's defaults, tools/dispersion.py:67-69)
GUESSES = (0.4739 + 0.153j, 1.793 + 0.491j, 0.9371 + 0.287j)

# the default input as Maxwellian components: bulk (n = 0.9, T = 1) + bump (1 - n = 0.1, v0 = 5, T2 = 1), electrons
BUMP_ON_TAIL = ((-1.0, 1.0, 1.0, 0.9, 0.0), (-1.0, 1.0, 1.0, 0.1, 5.0))
THERMAL = ((-1.0, 1.0, 1.0, 1.0, 0.0),)


def plasma_z(zeta: complex) -> complex:
    """Z(zeta) = i sqrt(pi) w(zeta), w = Faddeeva function."""
    return 1j * math.sqrt(math.pi) * wofz(zeta)


def dispersion_function(omega: complex, k: float, species: Sequence[Species]) -> complex:
    d = 1.0 + 0.0j
    for charge, mass, temperature, density, v0 in species:
        vth2 = temperature / mass
        zeta = (omega / k - v0) / math.sqrt(2.0 * vth2)
        d += density * charge ** 2 / mass / (k ** 2 * vth2) * (1.0 + zeta * plasma_z(zeta))
    return d


def muller(f: Callable[[complex], complex], x0: complex, x1: complex, x2: complex, ftol: float = 1e-14,
           xtol: float = 1e-14, max_iter: int = 100) -> complex:
    """Muller's method: fit a parabola through the last three iterates, step to its nearer root."""
    f0, f1, f2 = f(x0), f(x1), f(x2)
    for _ in range(max_iter):
        if abs(f2) <= ftol or abs(x2 - x1) <= xtol:
            break
        d01, d12, d02 = (f1 - f0) / (x1 - x0), (f2 - f1) / (x2 - x1), (f2 - f0) / (x2 - x0)
        w = d12 + d02 - d01
        curv = (d12 - d01) / (x2 - x0)
        root = cmath.sqrt(w * w - 4.0 * f2 * curv)
        den = w + root if abs(w + root) > abs(w - root) else w - root
        x0, x1, x2 = x1, x2, x2 - 2.0 * f2 / den
        f0, f1, f2 = f1, f2, f(x2)
    return x2


def solve_omega(k: float, species: Sequence[Species], guesses: Sequence[complex] = GUESSES) -> complex:
    return muller(lambda om: dispersion_function(om, k, species), *guesses)


if __name__ == "__main__":
    om = solve_omega(0.36, BUMP_ON_TAIL)
    print(f"bump-on-tail, k = 0.36: omega = {om.real:.7f} {om.imag:+.7f} i")
    om = solve_omega(0.5, THERMAL, (1.4 - 0.1j, 1.5 - 0.2j, 1.45 - 0.15j))
    print(f"thermal plasma, k = 0.5: omega = {om.real:.7f} {om.imag:+.7f} i")
