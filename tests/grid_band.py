"""TEST-ONLY stand-in for greb_b200.bigrid.DeviceBand: the same band / halo bookkeeping on host
memory with oracle/grid_oracle.c doing the arithmetic, so that the exchange logic of
greb_b200.bigrid.advance can run under gloo on a machine without GPUs."""
import ctypes as C

import numpy as np

from oracle import grid as og


class OracleBand:
    def __init__(self, nx, ny, k0, k1, s, pi=3.1416, kappa=8e5):
        self.nx, self.ny, self.k0, self.k1, self.s = nx, ny, k0, k1, s
        self.g = og.Geometry(nx, ny, pi, kappa)
        self.nsub = self.g.nsub
        self.kbase = max(0, k0 - 2 * s)
        self.nrows = min(ny, k1 + 2 * s) - self.kbase
        self.valid = (self.kbase, self.kbase + self.nrows)

    def set_fields(self, X, wz, u, v):
        cut = lambda a: np.ascontiguousarray(a[self.kbase:self.kbase + self.nrows], dtype=np.float32).copy()
        self.X = [cut(X), cut(X)]
        self.wz, self.u, self.v = cut(wz), cut(u), cut(v)
        self.cur = 0
        self.valid = (self.kbase, self.kbase + self.nrows)

    def _shift(self, a):
        return C.cast(a.ctypes.data - self.kbase * self.nx * 4, og.fp)     # index by GLOBAL row

    def substeps(self, n):
        g = self.g
        for _ in range(n):
            lo = self.valid[0] + 2 if self.valid[0] > 0 else 0
            hi = self.valid[1] - 2 if self.valid[1] < self.ny else self.ny
            assert lo <= self.k0 and hi >= self.k1, "halo used up"
            og.lib().gg_substep(self.nx, self.ny, lo, hi, self._shift(self.X[self.cur]), self._shift(self.wz),
                                self._shift(self.u), self._shift(self.v), g.ccy_diff, g.ccy_adv, og._p(g.ccx_diff),
                                og._p(g.ccx_adv), og._p(g.ccx2_diff), og._p(g.ccx2_adv), og._pi(g.polar),
                                og._pi(g.time2_diff), og._pi(g.time2_adv), self._shift(self.X[self.cur ^ 1]))
            self.cur ^= 1
            self.valid = (lo, hi)

    def substeps_async(self, n):
        self.substeps(n)

    def sync(self):
        pass

    def rows(self, lo, hi):
        import torch
        return torch.from_numpy(self.X[self.cur][lo - self.kbase:hi - self.kbase])

    def halo_refreshed(self):
        self.valid = (self.kbase, self.kbase + self.nrows)

    def get(self):
        return self.X[self.cur][self.k0 - self.kbase:self.k1 - self.kbase].copy()
