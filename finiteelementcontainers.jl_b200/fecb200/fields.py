"""Field containers (src/Fields.jl): H1Field and Connectivity with the reference's layouts."""
from __future__ import annotations

import numpy as np


class H1Field(np.ndarray):
    """H1Field{T,D,NF} (src/Fields.jl:237-259): a (NF, NN) view whose flat `data` is node-major /
    dof-fastest: data[(n-1)*NF + d] (src/Fields.jl:36-40).  Implemented as a Fortran-ordered
    ndarray subclass so `field[d, n]` and `field.data_flat` both follow the reference."""

    def __new__(cls, arr):
        a = np.asfortranarray(np.asarray(arr, dtype=np.float64))
        if a.ndim == 1:
            a = a.reshape(1, -1, order="F")
        return a.view(cls)

    @classmethod
    def zeros(cls, nf, nn):
        return cls(np.zeros((nf, nn), order="F"))

    @property
    def data_flat(self) -> np.ndarray:
        """the reference's `field.data` (flat, contiguous, shares memory)"""
        return np.asarray(self).reshape(-1, order="F")


class Connectivity:
    """Connectivity (src/Fields.jl:129-160): per-block connectivities concatenated into one flat
    1-based Int64 vector, element-major, with 1-based block `offsets`."""

    def __init__(self, mats):
        self.nblocks = len(mats)
        self.nepes = [m.shape[0] for m in mats]
        self.nelems = [m.shape[1] for m in mats]
        self.offsets = []
        off = 1
        for nepe, nel in zip(self.nepes, self.nelems):
            self.offsets.append(off)
            off += nepe * nel
        self.data = np.concatenate([np.asarray(m, dtype=np.int64).reshape(-1, order="F") for m in mats])

    def block(self, b):
        """connectivity(conn, b) (:163-168), 0-based b here; returns (NNPE, NE)"""
        o = self.offsets[b] - 1
        return self.data[o:o + self.nepes[b] * self.nelems[b]].reshape(self.nepes[b], self.nelems[b], order="F")
