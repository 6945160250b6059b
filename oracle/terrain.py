"""Piecewise terrain of the reference, CPU oracle (TEST INFRASTRUCTURE ONLY).

Restates `generate_piecewise_terrain` / `piecewise1_2D_lc`, `piecewise2_2D_lc`
(src/simulation/environments/piecewise.jl:40-132) and the surface rotation of a planar environment
(src/simulator/environment.jl:81-96).  Written for scalars that may be COMPLEX: `height` continues the surface along its
tangent, s(x + iδ) = s(x) + iδ s'(x), so that complex-step differentiation of a residual that contains the surface returns
the exact first derivative (oracle/residual.py::TerrainResidual)."""
from __future__ import annotations

import math

import numpy as np


class PiecewiseTerrain:
    def __init__(self, slope_deg: float = 10.0):
        self.m = m = math.tan(math.radians(slope_deg))  # piecewise.jl:123, 129
        # each kink: cubic through (x1, y1), (x2, y2) with end slopes m1, m2  (piecewise.jl:41-78)
        self.kinks = []
        for (x1, y1, m1, x2, y2, m2) in ((0.4, 0.0, 0.0, 0.6, m * 0.1, m),
                                         (1.4, m * 1.4, m, 1.6, m * 1.5 + (-0.25 * m) * 0.1, -0.25 * m)):
            rows = [[x ** 3, x ** 2, x, 1.0] for x in (x1, x2)] + [[3 * x ** 2, 2 * x, 1.0, 0.0] for x in (x1, x2)]
            self.kinks.append(np.linalg.solve(np.array(rows), np.array([y1, y2, m1, m2])))

    def _eval(self, x: float):
        m, (a, b) = self.m, self.kinks
        if x < 0.4:
            return 0.0, 0.0
        if x < 0.6:
            return np.polyval(a, x), np.polyval(np.polyder(a), x)
        if x < 1.9:
            return m * x - 0.5 * m, m
        if x < 2.1:  # the second cubic lives in a coordinate shifted by 0.5 (piecewise.jl:86, 94)
            return np.polyval(b, x - 0.5), np.polyval(np.polyder(b), x - 0.5)
        return -0.25 * m * (x - 2.0) + 1.5 * m, -0.25 * m

    def slope(self, x) -> float:
        return float(self._eval(float(np.real(x)))[1])

    def height(self, x):
        s, ds = self._eval(float(np.real(x)))
        return s + 1j * np.imag(x) * ds if np.iscomplexobj(x) else float(s)

    def rotation(self, x):
        """(cos, sin) of the world→surface rotation [[c, s], [−s, c]] under x (environment.jl:81-96): the surface normal
        n = (−s', 1)/‖·‖ is turned onto (0, 1), ang = atan2(1, 0) − atan2(n_y, n_x), R = [[cos ang, −sin ang], [sin ang, cos ang]]."""
        sg = self.slope(x)
        n = np.array([-sg, 1.0]) / math.hypot(sg, 1.0)
        ang = math.atan2(1.0, 0.0) - math.atan2(n[1], n[0])
        return math.cos(ang), -math.sin(ang)


def get_terrain(name: str) -> PiecewiseTerrain:
    return {"piecewise1_2D_lc": PiecewiseTerrain(10.0), "piecewise2_2D_lc": PiecewiseTerrain(-10.0)}[name]
