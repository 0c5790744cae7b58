"""Minimal NumPy stand-in for the QuTiP calls the reference makes on the CZ path.

FIXTURE-GENERATION TOOL (this container only).  QuTiP is not installable here
(SURVEY.md section 8c), so to run the reference's *own, unmodified* orchestration code
(``RG/simulation.py:simulate_CZ_gate`` and everything it calls) we register this
module as ``qutip`` before importing the reference.  Every QuTiP numeric is
answered by the oracle restatement in ``oracle/rydcz_oracle.py`` (``mesolve`` ->
matrix exponential or zvode); everything else -- the scalar prologue, protocol
logic, quirks, result packing -- is the reference's real code.

Only ``tests/golden/make_golden.py`` imports this file.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import scipy.linalg as sla

MESOLVE_MODE = "exact"  # or "zvode"
MESOLVE_CALLS = 0


class Qobj:
    __array_priority__ = 1000

    def __init__(self, data):
        self.data = np.array(data, dtype=complex)
        if self.data.ndim == 1:
            self.data = self.data.reshape(-1, 1)

    # -- type predicates -------------------------------------------------
    @property
    def isket(self):
        return self.data.shape[1] == 1 and self.data.shape[0] > 1

    @property
    def isbra(self):
        return self.data.shape[0] == 1 and self.data.shape[1] > 1

    @property
    def isoper(self):
        return self.data.shape[0] == self.data.shape[1] and self.data.shape[0] > 1

    @property
    def shape(self):
        return self.data.shape

    # -- algebra -----------------------------------------------------------
    def dag(self):
        return Qobj(self.data.conj().T)

    def full(self):
        return self.data.copy()

    def tr(self):
        return np.trace(self.data)

    def norm(self):
        return np.linalg.norm(self.data)

    def __mul__(self, other):
        if isinstance(other, Qobj):
            out = self.data @ other.data
            if out.shape == (1, 1):
                return Qobj(out)  # QuTiP 5 keeps 1x1 Qobj for bra*ket; callers use .full()
            return Qobj(out)
        return Qobj(self.data * other)

    def __rmul__(self, other):
        return Qobj(self.data * other)

    __matmul__ = __mul__

    def __add__(self, other):
        if isinstance(other, Qobj):
            return Qobj(self.data + other.data)
        if other == 0:
            return Qobj(self.data)
        return Qobj(self.data + other * np.eye(self.data.shape[0]))

    __radd__ = __add__

    def __sub__(self, other):
        return Qobj(self.data - (other.data if isinstance(other, Qobj) else other))

    def __neg__(self):
        return Qobj(-self.data)

    def __truediv__(self, s):
        return Qobj(self.data / s)

    def eigenstates(self):
        w, v = sla.eigh(self.data)
        kets = np.empty((v.shape[1],), dtype=object)
        kets[:] = [Qobj(v[:, i]) for i in range(v.shape[1])]
        return w, kets


def basis(n, i):
    v = np.zeros(n, dtype=complex)
    v[i] = 1.0
    return Qobj(v)


def qeye(n):
    return Qobj(np.eye(n))


def tensor(*ops):
    out = ops[0].data
    for o in ops[1:]:
        out = np.kron(out, o.data)
    return Qobj(out)


def ket2dm(psi):
    return Qobj(psi.data @ psi.data.conj().T)


def fidelity(a, b):
    """sqrt-fidelity; the reference only calls it with one pure argument
    (RG/simulation.py:220-222), for which F = sqrt(<t|rho|t>)."""
    A, B = a.data, b.data
    for X, Y in ((A, B), (B, A)):
        w, v = np.linalg.eigh(Y)
        if abs(w[-1] - 1.0) < 1e-12 and np.all(np.abs(w[:-1]) < 1e-12):
            t = v[:, -1]
            return float(np.sqrt(max(0.0, np.real(np.vdot(t, X @ t)))))
    sa = sla.sqrtm(A)
    ev = np.linalg.eigvals(sa @ B @ sa)
    return float(np.real(np.sum(np.sqrt(np.abs(ev)))))


class _Result:
    def __init__(self, states):
        self.states = states


def mesolve(H, psi0, tlist, c_ops=None, options=None, **kw):
    global MESOLVE_CALLS
    MESOLVE_CALLS += 1
    from oracle import rydcz_oracle as orc

    c_ops = c_ops or []
    Hm = H.data
    t_end = float(tlist[-1])
    n_t = len(tlist)
    nsteps = (options or {}).get("nsteps", 50000)
    if not c_ops and psi0.isket:
        gen, y0, shape = -1j * Hm, psi0.data[:, 0], None
    else:
        rho0 = psi0.data @ psi0.data.conj().T if psi0.isket else psi0.data
        gen, y0, shape = orc.liouvillian(Hm, [c.data for c in c_ops]), rho0.reshape(-1), rho0.shape
    if MESOLVE_MODE == "exact":
        y = sla.expm(gen * t_end) @ y0
    else:
        y = orc._zvode_advance(gen, y0, t_end, n_t, nsteps)
    out = Qobj(y if shape is None else y.reshape(shape))
    return _Result([psi0, out])


class Options(dict):
    pass


def install():
    mod = types.ModuleType("qutip")
    for name in ("Qobj", "basis", "qeye", "tensor", "ket2dm", "fidelity", "mesolve", "Options"):
        setattr(mod, name, globals()[name])
    sys.modules["qutip"] = mod
    return mod


def import_reference(root="/root/reference"):
    """Import the reference's rydberg_gates modules without running its
    ``__init__`` (which pulls in matplotlib).  Returns the stub package."""
    import importlib

    install()
    rg = f"{root}/src/qpu_simulator/micro_physics/neutral_atoms/rydberg_gates"
    pkg_name = "ref_rydberg_gates"
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = [rg]
    sys.modules[pkg_name] = pkg
    for sub in ("constants", "atom_database", "laser_physics", "trap_physics", "protocols",
                "pulse_shaping", "configurations", "noise_models", "hamiltonians", "simulation"):
        setattr(pkg, sub, importlib.import_module(f"{pkg_name}.{sub}"))
    return pkg
