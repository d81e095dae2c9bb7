"""A stand-in for libpolymc_b200.so with the C ABI's entry points as Python callables — test infrastructure for running
the Julia hosts (polymer-stats_b200/julia/*.jl) under tools/minijl WITHOUT a GPU.

It computes nothing physical: every output array is filled with a fixed formula of (chain, row, column), so a test can
predict what the host must print after ITS pooling / formatting / file writing.  What it does check for real is the
marshalling: `pmc_create` reads the case array through the package's own ctypes mirror of `pmc_case`
(polymc.lib.PmcCase), so a field the Julia struct puts in the wrong place or width shows up in `self.cases`.
Arguments arrive as the ctypes values minijl.ffi marshalled (c_void_p, c_int64, c_double …)."""
import ctypes as C

import numpy as np


def _v(x):
    return x.value if hasattr(x, "value") else x


def _arr(p, shape):
    n = int(np.prod(shape))
    buf = (C.c_double * n).from_address(_v(p))
    return np.frombuffer(buf, dtype=np.float64).reshape(shape)


def acc_value(chain, k):
    """Raw sum k (0..15) of a chain and its normaliser (k = 16)."""
    return 1000.0 * (chain + 1) if k == 16 else (k + 1) * 0.25 * (chain + 2) * (-1.0 if k % 3 == 1 else 1.0)


def extra_value(chain, k):
    return (k + 1) * 0.5 * (chain + 1)


def diag_value(chain, k):
    return float(137 * (chain + 1)) if k == 4 else float(k)


def traj_value(chain, row, k, step):
    return float(step) if k == 0 else (chain + 1) * 0.1 + row + 0.01 * k


def state_value(chain, row, k):
    return 0.3 + 0.001 * k + 0.01 * row + 0.1 * chain


class MockLib:
    def __init__(self, PmcCase, gpus=2):
        self.PmcCase = PmcCase
        self.gpus = gpus
        self.calls = []
        self.cases = []
        self.step = 0
        self.R = 0
        self.n = 0
        self.precision = None

    # ---- single device ------------------------------------------------------------------------------------------
    def pmc_last_error(self):
        return b"mock error"

    def pmc_create(self, cases, ncases, replicas, seed, device, base, out):
        arr = C.cast(_v(cases), C.POINTER(self.PmcCase))
        self.cases = [self.PmcCase.from_buffer_copy(arr[i]) for i in range(_v(ncases))]
        self.R, self.n, self.step = _v(replicas) * _v(ncases), int(self.cases[0].n), 0
        self.calls.append(("pmc_create", _v(ncases), _v(replicas), _v(seed), _v(device), _v(base)))
        C.cast(_v(out), C.POINTER(C.c_void_p))[0] = 0x1234
        return 0

    def pmc_multi_create(self, cases, ncases, replicas, seed, devices, ndev, out):
        rc = self.pmc_create(cases, ncases, replicas, seed, C.c_int32(-1), C.c_uint32(0), out)
        self.calls[-1] = ("pmc_multi_create", _v(ncases), _v(replicas), _v(seed), _v(devices), _v(ndev))
        return rc

    def pmc_set_pair_precision(self, h, mode):
        self._check(h)
        self.precision = _v(mode)
        return 0

    pmc_multi_set_pair_precision = pmc_set_pair_precision

    def _check(self, h):
        assert _v(h) == 0x1234, "host passed a wrong handle"

    def pmc_rows_for(self, h, todo, stepout):
        self._check(h)
        so = _v(stepout)
        return 0 if so <= 0 else (self.step + _v(todo)) // so - self.step // so

    pmc_multi_rows_for = pmc_rows_for

    def _run(self, name, h, todo, stepout, traj, roll, state, cols):
        self._check(h)
        rows = self.pmc_rows_for(h, todo, stepout)
        so = _v(stepout)
        self.calls.append((name, _v(todo), so, rows))
        first = (self.step // so + 1) * so if so > 0 else 0
        if rows > 0 and _v(traj):
            t, r = _arr(traj, (self.R, rows, 8)), _arr(roll, (self.R, rows, cols))
            for c in range(self.R):
                for w in range(rows):
                    st = first + w * so
                    t[c, w] = [traj_value(c, w, k, st) for k in range(8)]
                    r[c, w] = [float(st)] + [acc_value(c, k) / acc_value(c, 16) * (w + 1) for k in range(cols - 1)]
            if state is not None and _v(state):
                s = _arr(state, (self.R, rows, 2 * self.n))
                for c in range(self.R):
                    for w in range(rows):
                        s[c, w] = [state_value(c, w, k) for k in range(2 * self.n)]
        self.step += _v(todo)
        return 0

    def pmc_run(self, h, todo, stepout, traj, roll):
        return self._run("pmc_run", h, todo, stepout, traj, roll, None, 17)

    pmc_multi_run = pmc_run

    def pmc_run_ex(self, h, todo, stepout, traj, roll, state):
        return self._run("pmc_run_ex", h, todo, stepout, traj, roll, state, 19)

    def pmc_reinit(self, h, flags):
        self._check(h)
        self.calls.append(("pmc_reinit",))
        self.step = 0
        return 0

    def pmc_begin_stage(self, h, scale):
        self._check(h)
        self.calls.append(("pmc_begin_stage", _v(scale)))
        self.step = 0
        return 0

    def pmc_init_x0(self, h, x0, nx, dx0):
        self._check(h)
        k = _v(nx)
        self.calls.append(("pmc_init_x0", _arr(x0, (k,)).tolist(), _arr(dx0, (k,)).tolist()))
        return 0

    def pmc_accumulators(self, h, out):
        self._check(h)
        a = _arr(out, (self.R, 17))
        for c in range(self.R):
            a[c] = [acc_value(c, k) for k in range(17)]
        return 0

    def pmc_extra_accumulators(self, h, out):
        self._check(h)
        a = _arr(out, (self.R, 2))
        for c in range(self.R):
            a[c] = [extra_value(c, k) for k in range(2)]
        return 0

    def pmc_diagnostics(self, h, out):
        self._check(h)
        a = _arr(out, (self.R, 8))
        for c in range(self.R):
            a[c] = [diag_value(c, k) for k in range(8)]
        return 0

    def pmc_multi_gather(self, h, out):
        """[R][24]: 16 averages, acceptance rate, normaliser, …, column 20 = trials (include/polymc.h)."""
        self._check(h)
        a = _arr(out, (self.R, 24))
        for c in range(self.R):
            nrm = acc_value(c, 16)
            a[c, :16] = [acc_value(c, k) / nrm for k in range(16)]
            a[c, 16] = diag_value(c, 4) / 1000.0
            a[c, 17] = nrm
            a[c, 18:] = 0.0
            a[c, 20] = 1000.0
        self.calls.append(("pmc_multi_gather",))
        return 0

    def pmc_destroy(self, h):
        self._check(h)
        self.calls.append(("pmc_destroy",))

    pmc_multi_destroy = pmc_destroy
