"""Just enough of the petsc4py surface to drive ``ssc.PatchPC`` where PETSc is not installed.

The reference is used as ``-pc_type python -pc_python_type ssc.PatchPC`` under a PETSc KSP
(tests/test_jacobi_equivalent.py:66-76).  petsc4py/Firedrake are absent from this build's containers, so the
tests and the benchmark drive the same class through these duck-typed stand-ins: an options database with
prefixes, Vec/Mat wrappers over numpy/scipy, a PC of type "python" with firedrake.PCBase's
initialize-once-then-update dispatch, and CG/GMRES with PETSc's defaults (left preconditioning, preconditioned
residual norm, rtol 1e-5, atol 1e-50, dtol 1e5, max_it 1e4).  With a real petsc4py the same ``PatchPC`` methods
are called by PETSc itself; nothing here is on the GPU path.
"""
import importlib

import numpy as np


class Options(object):
    """PETSc.Options(prefix): one global database; keys are stored without the leading dash."""
    _db = {}

    def __init__(self, prefix=None):
        self.prefix = prefix or ""

    @classmethod
    def clear(cls):
        cls._db.clear()

    def _key(self, name):
        name = name[1:] if name.startswith("-") else name
        return self.prefix + name

    def setValue(self, name, value):
        if isinstance(value, bool):
            value = "true" if value else "false"
        Options._db[self._key(name)] = str(value)

    __setitem__ = setValue

    def delValue(self, name):
        Options._db.pop(self._key(name), None)

    def hasName(self, name):
        return self._key(name) in Options._db

    def getString(self, name, default=None):
        return Options._db.get(self._key(name), default)

    def getBool(self, name, default=None):
        v = self.getString(name)
        if v is None:
            return default
        return v.lower() in ("", "1", "true", "yes", "on")

    def getInt(self, name, default=None):
        v = self.getString(name)
        return default if v is None else int(v)

    def getAll(self):
        return dict(Options._db)

    @classmethod
    def insert_parameters(cls, params, prefix=""):
        """Flatten a Firedrake-style solver_parameters dict (nested dicts extend the key with '_')."""
        for k, v in params.items():
            if isinstance(v, dict):
                cls.insert_parameters(v, prefix + k + "_")
            else:
                Options(prefix).setValue(k, v)


class Vec(object):
    def __init__(self, array):
        self.array = np.ascontiguousarray(array, dtype=np.float64)

    @property
    def array_r(self):
        return self.array

    def getArray(self, readonly=False):
        return self.array

    def getLocalSize(self):
        return self.array.size

    getSize = getLocalSize

    def duplicate(self):
        return Vec(np.zeros_like(self.array))

    def copy(self):
        return Vec(self.array.copy())

    def set(self, v):
        self.array[:] = v


class Mat(object):
    """Assembled ("aij", scipy CSR inside) or shell ("python") matrix."""

    def __init__(self, csr=None, context=None, size=None):
        self.csr = csr
        self.context = context
        self._size = size if size is not None else (csr.shape if csr is not None else None)

    def getType(self):
        return "python" if self.csr is None else "seqaij"

    def getPythonContext(self):
        return self.context

    def getSize(self):
        return self._size

    def getValuesCSR(self):
        return self.csr.indptr, self.csr.indices, self.csr.data

    def mult(self, x, y):
        if self.csr is not None:
            y.array[:] = self.csr @ x.array
        else:
            self.context.mult(self, x, y)

    def createVecs(self):
        n = self._size[0]
        return Vec(np.zeros(n)), Vec(np.zeros(n))

    def setNearNullSpace(self, nullspace):
        self._nearnullsp = nullspace

    def getNearNullSpace(self):
        return getattr(self, "_nearnullsp", None)


class NullSpace(object):
    """petsc4py's NullSpace as far as ssc/lo.py:146-159 uses it: a bag of vectors."""

    def __init__(self, vectors=()):
        self._vecs = [v if isinstance(v, Vec) else Vec(np.asarray(v, dtype=np.float64)) for v in vectors]

    def create(self, vectors=(), comm=None):
        self.__init__(vectors)
        return self

    def getVecs(self):
        return list(self._vecs)


class DM(object):
    def __init__(self, appctx=None):
        self.appctx = appctx


def get_appctx(dm):
    return None if dm is None else dm.appctx


class PCBase(object):
    """firedrake.PCBase: setUp = initialize on the first call, update afterwards (external to the reference tree)."""

    def __init__(self):
        self.initialized = False

    def setUp(self, pc):
        if self.initialized:
            self.update(pc)
        else:
            self.initialize(pc)
            self.initialized = True
            self.update(pc)      # PatchPC.update is what runs PCSetUp_PATCH (ssc/patch.py:269-270)

    def view(self, pc, viewer=None):
        pass


class Viewer(object):
    def __init__(self):
        self.lines = []

    def printfASCII(self, s):
        self.lines.append(s)

    def __str__(self):
        return "".join(self.lines)


class PC(object):
    """PETSc.PC of type "jacobi", "none" or "python"."""

    def __init__(self, comm=None):
        self.comm = comm
        self.prefix = ""
        self.type = "none"
        self.A = self.P = None
        self.dm = None
        self.ctx = None
        self.attrs = {}
        self._setup = False

    def create(self, comm=None):
        self.comm = comm
        return self

    def setOptionsPrefix(self, prefix):
        self.prefix = prefix or ""

    def getOptionsPrefix(self):
        return self.prefix

    def setOperators(self, A, P=None):
        self.A, self.P = A, (A if P is None else P)
        self._setup = False

    def getOperators(self):
        return self.A, self.P

    def setDM(self, dm):
        self.dm = dm

    def getDM(self):
        return self.dm

    def setType(self, t):
        self.type = t

    def setPythonContext(self, ctx):
        self.ctx = ctx

    def getPythonContext(self):
        return self.ctx

    def setAttr(self, k, v):
        self.attrs[k] = v

    def getAttr(self, k):
        return self.attrs.get(k)

    def incrementTabLevel(self, n, parent=None):
        pass

    def setFromOptions(self):
        opts = Options(self.prefix)
        t = opts.getString("pc_type")
        if t is not None:
            self.type = t
        if self.type == "python" and self.ctx is None:
            name = opts.getString("pc_python_type")
            if name is None:
                raise ValueError("pc_type python needs %spc_python_type" % self.prefix)
            modname, clsname = name.rsplit(".", 1)
            self.ctx = getattr(importlib.import_module(modname), clsname)()

    def setUp(self):
        if self.type == "python":
            self.ctx.setUp(self)
        elif self.type == "jacobi":
            d = self.P.csr.diagonal()
            self._dinv = 1.0 / d
        self._setup = True

    def apply(self, x, y):
        if not self._setup:
            self.setUp()
        if self.type == "python":
            self.ctx.apply(self, x, y)
        elif self.type == "jacobi":
            y.array[:] = self._dinv * x.array
        else:
            y.array[:] = x.array

    def view(self, viewer=None):
        if self.type == "python":
            self.ctx.view(self, viewer)


class KSP(object):
    """cg / gmres / preonly with PETSc's default tolerances and norm type."""

    def __init__(self, comm=None):
        self.prefix = ""
        self.type = "gmres"
        self.pc = PC()
        self.rtol, self.atol, self.divtol, self.max_it = 1e-5, 1e-50, 1e5, 10000
        self.restart = 30
        self.history = None
        self.its = 0
        self.reason = 0
        self.dm = None

    def create(self, comm=None):
        return self

    def setOptionsPrefix(self, prefix):
        self.prefix = prefix or ""
        self.pc.setOptionsPrefix(self.prefix)

    def setDM(self, dm):
        self.dm = dm
        self.pc.setDM(dm)

    def setOperators(self, A, P=None):
        self.A = A
        self.pc.setOperators(A, P)

    def getPC(self):
        return self.pc

    def setFromOptions(self):
        o = Options(self.prefix)
        self.type = o.getString("ksp_type", self.type)
        self.rtol = float(o.getString("ksp_rtol", self.rtol))
        self.atol = float(o.getString("ksp_atol", self.atol))
        self.max_it = int(o.getString("ksp_max_it", self.max_it))
        self.restart = int(o.getString("ksp_gmres_restart", self.restart))
        self.pc.setFromOptions()

    def setConvergenceHistory(self):
        self.history = []

    def getConvergenceHistory(self):
        return np.array(self.history)

    def getIterationNumber(self):
        return self.its

    def _converged(self, rnorm, r0):
        if self.history is not None:
            self.history.append(rnorm)
        if rnorm != rnorm:
            self.reason = -9
            return True
        if rnorm <= max(self.rtol * r0, self.atol):
            self.reason = 2 if rnorm > self.atol else 3
            return True
        if rnorm >= self.divtol * r0:
            self.reason = -4
            return True
        return False

    def solve(self, b, x):
        self.pc.setUp()
        if self.history is not None:
            self.history = []
        if self.type == "cg":
            return self._cg(b, x)
        if self.type == "gmres":
            return self._gmres(b, x)
        if self.type == "preonly":
            self.pc.apply(b, x)
            self.its = 1
            return
        raise ValueError("unknown ksp_type %s" % self.type)

    def _cg(self, b, x):
        A, M = self.A, self.pc
        r = Vec(b.array.copy())
        tmp = r.duplicate()
        if np.any(x.array):
            A.mult(x, tmp)
            r.array -= tmp.array
        z = r.duplicate()
        M.apply(r, z)
        rz = float(r.array @ z.array)
        r0 = np.sqrt(abs(float(z.array @ z.array)))       # KSP_NORM_PRECONDITIONED
        self.its = 0
        self.reason = 0
        if self._converged(r0, r0):
            return
        p = Vec(z.array.copy())
        for it in range(1, self.max_it + 1):
            A.mult(p, tmp)
            alpha = rz / float(p.array @ tmp.array)
            x.array += alpha * p.array
            r.array -= alpha * tmp.array
            M.apply(r, z)
            rz_new = float(r.array @ z.array)
            self.its = it
            if self._converged(np.sqrt(abs(float(z.array @ z.array))), r0):
                return
            p.array[:] = z.array + (rz_new / rz) * p.array
            rz = rz_new
        self.reason = -3

    def _gmres(self, b, x):
        A, M = self.A, self.pc
        n = b.array.size
        self.its = 0
        self.reason = 0
        tmp, w = Vec(np.zeros(n)), Vec(np.zeros(n))
        r0 = None
        while True:
            A.mult(x, tmp)
            tmp.array[:] = b.array - tmp.array
            M.apply(tmp, w)                                  # left preconditioning
            beta = np.linalg.norm(w.array)
            if r0 is None:
                r0 = beta
                if self._converged(beta, r0):
                    return
            m = self.restart
            V = np.zeros((m + 1, n))
            H = np.zeros((m + 1, m))
            cs, sn, g = np.zeros(m), np.zeros(m), np.zeros(m + 1)
            V[0] = w.array / beta
            g[0] = beta
            k_done = 0
            done = False
            for k in range(m):
                A.mult(Vec(V[k]), tmp)
                M.apply(tmp, w)
                for i in range(k + 1):                       # classical Gram-Schmidt with one refinement is PETSc's default; MGS is equivalent here
                    H[i, k] = w.array @ V[i]
                    w.array -= H[i, k] * V[i]
                H[k + 1, k] = np.linalg.norm(w.array)
                if H[k + 1, k] != 0:
                    V[k + 1] = w.array / H[k + 1, k]
                for i in range(k):
                    t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                    H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                    H[i, k] = t
                d = np.hypot(H[k, k], H[k + 1, k])
                cs[k], sn[k] = H[k, k] / d, H[k + 1, k] / d
                H[k, k] = d
                H[k + 1, k] = 0
                g[k + 1] = -sn[k] * g[k]
                g[k] = cs[k] * g[k]
                self.its += 1
                k_done = k + 1
                if self._converged(abs(g[k + 1]), r0) or self.its >= self.max_it:
                    done = True
                    break
            y = np.linalg.solve(np.triu(H[:k_done, :k_done]), g[:k_done])
            x.array += V[:k_done].T @ y
            if done:
                return
