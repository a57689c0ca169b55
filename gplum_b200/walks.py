"""Container for one FDPS force pass in index form (what the multi-walk-index interface sees,
FDPS/src/tree_for_force_impl_force.hpp:232-238): i-groups + index lists into epj_all / spj_all."""
import numpy as np

from . import structs as S


class Walks:
    """One FDPS force pass in index form: i-groups + index lists into epj_all / spj_all."""

    def __init__(self, epi, epi_off, ni, adr_epj, epj_disp, n_epj, adr_spj, spj_disp, n_spj, epj_all, spj_all):
        self.epi = np.ascontiguousarray(epi, dtype=S.EPI)
        self.epi_off = np.ascontiguousarray(epi_off, dtype=np.int32)
        self.ni = np.ascontiguousarray(ni, dtype=np.int32)
        self.adr_epj = np.ascontiguousarray(adr_epj, dtype=np.int32)
        self.epj_disp = np.ascontiguousarray(epj_disp, dtype=np.int64)
        self.n_epj = np.ascontiguousarray(n_epj, dtype=np.int32)
        self.adr_spj = np.ascontiguousarray(adr_spj, dtype=np.int32)
        self.spj_disp = np.ascontiguousarray(spj_disp, dtype=np.int64)
        self.n_spj = np.ascontiguousarray(n_spj, dtype=np.int32)
        self.epj_all = np.ascontiguousarray(epj_all, dtype=S.EPJ)
        self.spj_all = np.ascontiguousarray(spj_all)

    @property
    def n_walk(self):
        return len(self.ni)

    @property
    def quad(self):
        return self.spj_all.dtype.itemsize == 80

    def n_interactions(self):
        ni = self.ni.astype(np.int64)
        return int((ni * self.n_epj).sum()), int((ni * self.n_spj).sum())

    def save(self, path):
        np.savez_compressed(path, **{k: getattr(self, k) for k in
                                     ("epi", "epi_off", "ni", "adr_epj", "epj_disp", "n_epj", "adr_spj",
                                      "spj_disp", "n_spj", "epj_all", "spj_all")})

    @classmethod
    def load(cls, path):
        z = np.load(path)
        return cls(*[z[k] for k in ("epi", "epi_off", "ni", "adr_epj", "epj_disp", "n_epj", "adr_spj",
                                    "spj_disp", "n_spj", "epj_all", "spj_all")])
