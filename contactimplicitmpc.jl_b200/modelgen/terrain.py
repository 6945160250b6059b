"""Non-flat 2-D terrain of the reference: `piecewise1_2D_lc` / `piecewise2_2D_lc`
(src/simulation/environments/piecewise.jl:40-132): flat, a 10° ramp from x = 0.5, a −2.5° descent from x = 2.0, the
two kinks smoothed by cubics over ±0.1 m.

The cubic coefficients are the solutions of the two 4×4 interpolation systems of `generate_piecewise_terrain`
(:41-78): value and slope matched at both ends of a kink.  The second cubic is fitted in a coordinate shifted by 0.5
(x1 = 1.4, x2 = 1.6 for the kink at 1.9 … 2.1) and evaluated at `x − 0.5` (:86, :94) — kept as is.

`Terrain.surf / dsurf` evaluate height and slope on the host (numpy); `Terrain.c_source()` is the text the code
generator puts into `csrc/gen/residual_<robot>_piecewise.h` for the device."""
from __future__ import annotations

import math

import numpy as np


def _cubic(x1, y1, m1, x2, y2, m2):
    A = np.array([[x1 ** 3, x1 ** 2, x1, 1.0], [x2 ** 3, x2 ** 2, x2, 1.0],
                  [3 * x1 ** 2, 2 * x1, 1.0, 0.0], [3 * x2 ** 2, 2 * x2, 1.0, 0.0]])
    return np.linalg.solve(A, np.array([y1, y2, m1, m2]))


class Terrain:
    """height s(x) and slope s'(x) of one piecewise environment; `slope_deg` = +10 (piecewise1) or −10 (piecewise2)."""

    def __init__(self, name: str, slope_deg: float):
        self.name = name
        m = math.tan(math.radians(slope_deg))
        self.m = m
        self.a1 = _cubic(0.4, 0.0, 0.0, 0.6, 0.1 * m, m)                                  # :43-58
        self.a2 = _cubic(1.4, 1.4 * m, m, 1.6, 1.5 * m - 0.25 * m * 0.1, -0.25 * m)       # :61-78
        # the reference asserts the fit at the ends of each kink (:57-58, :77-78)
        assert abs(self._poly(self.a1, 0.6) - 0.1 * m) < 1e-8 and abs(self._poly(self.a2, 1.4) - 1.4 * m) < 1e-8

    @staticmethod
    def _poly(a, x):
        return a[3] + a[2] * x + a[1] * x ** 2 + a[0] * x ** 3

    @staticmethod
    def _dpoly(a, x):
        return a[2] + 2.0 * a[1] * x + 3.0 * a[0] * x ** 2

    def surf(self, x: float) -> float:  # piecewise_smoothed, :81-88
        m = self.m
        if x < 0.4:
            return 0.0
        if x < 0.6:
            return float(self._poly(self.a1, x))
        if x < 1.9:
            return m * x - 0.5 * m
        if x < 2.1:
            return float(self._poly(self.a2, x - 0.5))
        return -0.25 * m * (x - 2.0) + 1.5 * m

    def dsurf(self, x: float) -> float:  # d_piecewise_smoothed, :90-96
        m = self.m
        if x < 0.4:
            return 0.0
        if x < 0.6:
            return float(self._dpoly(self.a1, x))
        if x < 1.9:
            return m
        if x < 2.1:
            return float(self._dpoly(self.a2, x - 0.5))
        return -0.25 * m

    def c_source(self) -> str:
        a1, a2, m = self.a1, self.a2, self.m
        f = lambda v: repr(float(v))  # noqa: E731  (shortest round-trip decimal)
        return "\n".join([
            f"// terrain `{self.name}` (src/simulation/environments/piecewise.jl:81-96): height s and slope ds at x",
            "CIMPC_GEN_HD void surf(const double x, double& s, double& ds) {",
            f"  const double m = {f(m)};",
            "  if (x < 0.4) { s = 0.0; ds = 0.0; }",
            f"  else if (x < 0.6) {{ s = {f(a1[3])} + x * ({f(a1[2])} + x * ({f(a1[1])} + x * {f(a1[0])}));",
            f"    ds = {f(a1[2])} + x * ({f(2 * a1[1])} + x * {f(3 * a1[0])}); }}",
            "  else if (x < 1.9) { s = m * x - 0.5 * m; ds = m; }",
            f"  else if (x < 2.1) {{ const double y = x - 0.5; s = {f(a2[3])} + y * ({f(a2[2])} + y * ({f(a2[1])} + y * {f(a2[0])}));",
            f"    ds = {f(a2[2])} + y * ({f(2 * a2[1])} + y * {f(3 * a2[0])}); }}",
            "  else { s = -0.25 * m * (x - 2.0) + 1.5 * m; ds = -0.25 * m; }",
            "}"])


TERRAINS = {"piecewise1_2D_lc": lambda: Terrain("piecewise1_2D_lc", 10.0),
            "piecewise2_2D_lc": lambda: Terrain("piecewise2_2D_lc", -10.0)}


def get_terrain(name: str) -> Terrain:
    return TERRAINS[name]()
