"""numpy views of the reference's particle / force structs (default macro set:
USE_INDIVIDUAL_CUTOFF, USE_QUAD, cartesian coordinates).

Layouts follow /root/reference/src/particle.h:16-24 (NeighborInfo), :70-86 (ForceGrav),
:93-109 (EPIGrav), :149-156 (EPJGrav) and FDPS/src/tree.hpp:883-967 (SPJMonopole /
SPJQuadrupole; quad order xx,yy,zz,xy,xz,yz per FDPS/src/matrix_sym3.hpp:12).  The sizes and
offsets are pinned against the compiled reference by tests/test_oracle_vs_ref.py (ref_layout).
"""
import numpy as np

EPI = np.dtype([("id_local", "<i4"), ("myrank", "<i4"), ("pos", "<f8", (3,)),
                ("r_out", "<f8"), ("r_search", "<f8")], align=True)
EPJ = np.dtype([("id_local", "<i4"), ("myrank", "<i4"), ("pos", "<f8", (3,)),
                ("r_out", "<f8"), ("r_search", "<f8"), ("id", "<i8"), ("mass", "<f8"),
                ("vel", "<f8", (3,)), ("acc_d", "<f8", (3,))], align=True)
SPJ_QUAD = np.dtype([("mass", "<f8"), ("pos", "<f8", (3,)), ("quad", "<f8", (6,))], align=True)
SPJ_MONO = np.dtype([("mass", "<f8"), ("pos", "<f8", (3,))], align=True)
FORCE = np.dtype([("acc", "<f4", (3,)), ("phi", "<f4"), ("number", "<i4"), ("rank", "<i4"),
                  ("id_max", "<i4"), ("id_min", "<i4")], align=True)

assert EPI.itemsize == 48 and EPJ.itemsize == 112 and SPJ_QUAD.itemsize == 80
assert SPJ_MONO.itemsize == 32 and FORCE.itemsize == 32

ID_MIN_CLEAR = 2147483647   # S32_MAX-1 with S32_MAX = 1LL<<31 (src/main_p3t.cpp:32, particle.h:51)
ID_MAX_CLEAR = -1


def cleared_force(n):
    """ForceGrav::clear() applied to n entries (src/particle.h:81-85)."""
    f = np.zeros(n, dtype=FORCE)
    f["id_max"] = ID_MAX_CLEAR
    f["id_min"] = ID_MIN_CLEAR
    return f

# ---- changeover correction (correctForceLong, src/gravity_soft.h:245-372): results per i-particle ----
# CORR      acc/phi are the FP64 corrections to ADD to the (widened) tree force; `number` neighbours
#           start at ngb[ngb_off]; id_local is the pp index FDPS's write-back would target
# CORR_INIT the extra sums of correctForceLongInitial (:375-528)
# NGB       NeighborId (src/neighbor.h:205-245)
CORR = np.dtype([("acc", "<f8", (3,)), ("phi", "<f8"), ("acc0", "<f8"), ("id_cluster", "<i8"),
                 ("number", "<i4"), ("id_local", "<i4"), ("ngb_off", "<i4"), ("in_domain", "<i4")], align=True)
CORR_INIT = np.dtype([("acc_d", "<f8", (3,)), ("jerk_d", "<f8", (3,)), ("phi_d", "<f8"), ("pad", "<f8")], align=True)
NGB = np.dtype([("id", "<i8"), ("rank", "<i4"), ("id_local", "<i4")], align=True)
CORR_PARAMS = np.dtype([("eps2", "<f8"), ("dt_tree", "<f8"), ("gamma", "<f8"), ("R_search2", "<f8"),
                        ("R_search3", "<f8"), ("re_search", "<i4"), ("initial", "<i4")], align=True)
assert CORR.itemsize == 64 and CORR_INIT.itemsize == 64 and NGB.itemsize == 16 and CORR_PARAMS.itemsize == 48


def corr_params(eps2=0.0, dt_tree=2.0 ** -6, gamma=0.5, R_search2=1.0, R_search3=4.0, re_search=True, initial=False):
    """Defaults: sample/parameter.dat (dt_tree, gamma) and src/particle.h:927-928 (R_search2/3);
    re_search mirrors `#define USE_RE_SEARCH_NEIGHBOR` (src/main_p3t.cpp:15)."""
    p = np.zeros(1, dtype=CORR_PARAMS)
    p["eps2"], p["dt_tree"], p["gamma"], p["R_search2"], p["R_search3"] = eps2, dt_tree, gamma, R_search2, R_search3
    p["re_search"], p["initial"] = int(re_search), int(initial)
    return p
