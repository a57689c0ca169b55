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
