"""ctypes wrapper of oracle/_build/libfec_oracle.so (fec_oracle_c.c) -- TEST INFRASTRUCTURE /
CPU BASELINE ONLY (see the header of fec_oracle_c.c).  Build with `make -C oracle`."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libfec_oracle.so")
_lib = None

PHYS = {"poisson": 1, "linear": 2, "neo": 3, "neo_as_written": 4, "j2": 5}
_i64p, _f64p = C.POINTER(C.c_int64), C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run `make -C oracle`")
        _lib = C.CDLL(LIB_PATH)
        _lib.fec_oracle_sparse_csc.restype = C.c_int64
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


class CProblem:
    """One element block in the reference's layouts: conn (NNPE,NE) 1-based, X (ND,NN), tables."""

    def __init__(self, conn, X, tables, phys, nf, props=(), source_q=None, state_old=None):
        self.nnpe, self.ne = conn.shape
        self.nd, self.nn = X.shape
        self.nf = nf
        self.conn = np.ascontiguousarray(conn.T, dtype=np.int64).reshape(-1)      # element-major
        self.X = np.ascontiguousarray(X.T, dtype=float).reshape(-1)               # node-major
        N, dN, w = tables
        self.nq = len(w)
        self.N, self.dN, self.w = (np.ascontiguousarray(N, dtype=float), np.ascontiguousarray(dN, dtype=float),
                                   np.ascontiguousarray(w, dtype=float))
        self.phys = PHYS[phys]
        self.props = np.ascontiguousarray(np.asarray(props, dtype=float).reshape(-1)) if len(props) else np.zeros(1)
        self.source_q = None if source_q is None else np.ascontiguousarray(source_q, dtype=float).reshape(-1)
        ns = 7 if phys == "j2" else 0
        # reference layout [NS,NQ,NE] column-major == C-order (NE,NQ,NS)
        self.state_old = np.zeros(self.ne * self.nq * ns) if state_old is None else \
            np.ascontiguousarray(np.transpose(state_old, (2, 1, 0))).reshape(-1)
        self.state_new = np.zeros(self.ne * self.nq * ns)
        self.ns = ns

    def _args(self):
        return (C.c_int(self.nd), C.c_int(self.nnpe), C.c_int(self.nf), C.c_int(self.nq), C.c_int(self.phys),
                C.c_int64(self.ne), _p(self.conn, _i64p), _p(self.X, _f64p), _p(self.N, _f64p), _p(self.dN, _f64p),
                _p(self.w, _f64p), _p(self.props, _f64p), _p(self.source_q, _f64p), _p(self.state_old, _f64p),
                _p(self.state_new, _f64p))

    def assemble_vector(self, U_flat, nthreads=1):
        U = np.ascontiguousarray(U_flat, dtype=float)
        R = np.empty(self.nf * self.nn)
        lib().fec_oracle_assemble_vector(*self._args(), _p(U, _f64p), C.c_int64(len(R)), _p(R, _f64p), C.c_int(nthreads))
        return R

    def assemble_matrix_coo(self, U_flat, kind=2, nthreads=1, out=None):
        U = np.ascontiguousarray(U_flat, dtype=float)
        nd = self.nnpe * self.nf
        coo = np.empty(self.ne * nd * nd) if out is None else out
        lib().fec_oracle_assemble_matrix_coo(*self._args(), _p(U, _f64p), C.c_int(kind), _p(coo, _f64p), C.c_int(nthreads))
        return coo

    def assemble_action(self, U_flat, V_flat, kind=2, nthreads=1):
        U = np.ascontiguousarray(U_flat, dtype=float)
        V = np.ascontiguousarray(V_flat, dtype=float)
        out = np.empty(self.nf * self.nn)
        lib().fec_oracle_assemble_action(*self._args(), _p(U, _f64p), _p(V, _f64p), C.c_int(kind), C.c_int64(len(out)),
                                         _p(out, _f64p), C.c_int(nthreads))
        return out

    def pattern(self):
        nd = self.nnpe * self.nf
        Is = np.empty(self.ne * nd * nd, dtype=np.int64)
        Js = np.empty_like(Is)
        lib().fec_oracle_pattern(C.c_int(self.nnpe), C.c_int(self.nf), C.c_int64(self.ne), _p(self.conn, _i64p),
                                 _p(Is, _i64p), _p(Js, _i64p))
        return Is, Js

    def state_new_ref(self):
        return np.transpose(self.state_new.reshape(self.ne, self.nq, self.ns), (2, 1, 0))


class SparseWorkspace:
    """The cached arrays of SparseMatrixPattern (klasttouch, csrrowptr, csrcolval, csrnzval, ...)."""

    def __init__(self, ncoo, n):
        self.n, self.ncoo = n, ncoo
        self.klasttouch = np.zeros(n, dtype=np.int64)
        self.csrrowptr = np.zeros(n + 1, dtype=np.int64)
        self.csrcolval = np.zeros(ncoo, dtype=np.int64)
        self.csrnzval = np.zeros(ncoo)
        self.colptr = np.zeros(n + 1, dtype=np.int64)
        self.rowval = np.zeros(ncoo, dtype=np.int64)
        self.nzval = np.zeros(ncoo)

    def sparse_csc(self, Is, Js, slots, coo):
        nnz = lib().fec_oracle_sparse_csc(C.c_int64(len(Is)), _p(Is, _i64p), _p(Js, _i64p), _p(slots, _i64p),
                                          _p(coo, _f64p), C.c_int64(self.n), _p(self.klasttouch, _i64p),
                                          _p(self.csrrowptr, _i64p), _p(self.csrcolval, _i64p), _p(self.csrnzval, _f64p),
                                          _p(self.colptr, _i64p), _p(self.rowval, _i64p), _p(self.nzval, _f64p))
        return self.colptr, self.rowval[:nnz], self.nzval[:nnz]

    def csr(self, colptr, rowval, nzval):
        nnz = len(rowval)
        rowptr = np.zeros(self.n + 1, dtype=np.int64)
        colval = np.zeros(nnz, dtype=np.int64)
        nzr = np.zeros(nnz)
        lib().fec_oracle_csc_to_csr(C.c_int64(self.n), _p(colptr, _i64p), _p(np.ascontiguousarray(rowval), _i64p),
                                    _p(np.ascontiguousarray(nzval), _f64p), _p(rowptr, _i64p), _p(colval, _i64p),
                                    _p(nzr, _f64p))
        return rowptr, colval, nzr


def max_threads():
    return int(lib().fec_oracle_max_threads())
