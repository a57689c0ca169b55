"""Snapshot / restart wire formats of the reference, so that a run whose particles live in
HBM (gplum_b200.state) can start from and hand back the reference's files.

  snapNNNNNN.dat  ASCII: FileHeader::writeAscii (src/energy.h:100-110) then one FPGrav::writeAscii line per
                  particle (src/particle.h:844-858):  id mass r_planet f pos[3] vel[3] neighbor.number flag
  snap_tmp.dat    binary restart file: fwrite of FileHeader (src/energy.h:120-127; 128 B) followed by fwrite of
                  every FPGrav (src/particle.h:868-876; 344 B with the default macro set), written by
                  makeSnapTmp (src/func.h:67-80) and read back at src/main_p3t.cpp:275.

The record layouts below were measured from the compiled reference (sizeof / offsetof of FileHeader and FPGrav)
and are pinned by tests/test_snapshot.py.  FPGrav derives from EPJGrav, so the first 112 bytes of a record
ARE the EPJGrav the force path and the resident state use.
"""
import numpy as np

from . import structs as S

ENERGY_FIELDS = ("etot", "ekin", "ephi_sun", "ephi_planet", "ephi", "ephi_d", "edisp")
ENERGY = np.dtype([(k, "<f8") for k in ENERGY_FIELDS])
HEADER = np.dtype({"names": ["n_body", "id_next", "time", "e_init", "e_now"],
                   "formats": ["<i4", "<i4", "<f8", ENERGY, ENERGY],
                   "offsets": [0, 4, 8, 16, 72], "itemsize": 128})
_V = ("<f8", (3,))
NEIGHBOR = np.dtype([("number", "<i4"), ("rank", "<i4"), ("id_max", "<i4"), ("id_min", "<i4")])
FP = np.dtype({
    "names": ["id_local", "myrank", "pos", "r_out", "r_search", "id", "mass", "vel", "acc_d",
              "acc", "acc_s", "jerk_d", "jerk_s", "acc_gd", "phi", "phi_d", "phi_s", "v_disp", "r_out_inv",
              "time", "dt", "acc0", "r_planet", "f", "neighbor", "id_cluster", "n_cluster",
              "inDomain", "isSent", "isDead", "isMerged"],
    "formats": ["<i4", "<i4", _V, "<f8", "<f8", "<i8", "<f8", _V, _V,
                _V, _V, _V, _V, _V, "<f8", "<f8", "<f8", "<f8", "<f8",
                "<f8", "<f8", "<f8", "<f8", "<f8", NEIGHBOR, "<i8", "<i4",
                "u1", "u1", "u1", "u1"],
    "offsets": [0, 4, 8, 32, 40, 48, 56, 64, 88,
                112, 136, 160, 184, 208, 232, 240, 248, 256, 264,
                272, 280, 288, 296, 304, 312, 328, 336,
                340, 341, 342, 343],
    "itemsize": 344})
assert ENERGY.itemsize == 56 and HEADER.itemsize == 128 and FP.itemsize == 344
# what the ASCII line carries
ASCII = np.dtype([("id", "<i8"), ("mass", "<f8"), ("r_planet", "<f8"), ("f", "<f8"), ("pos", "<f8", (3,)),
                  ("vel", "<f8", (3,)), ("n_neighbor", "<i4"), ("flag", "<i4")])
_E_ASCII = ("etot", "ekin", "ephi_sun", "ephi_planet", "edisp")      # the 5 of 7 energies the ASCII header prints


# ------------------------------------------------------------------ binary restart file
def read_binary(path, max_particles=None):
    """(header[1], FP[n]) of a snap_tmp.dat.  max_particles reads a truncated file (test fixtures)."""
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) < HEADER.itemsize:
        raise ValueError("%s: shorter than a FileHeader" % path)
    header = np.frombuffer(raw, dtype=HEADER, count=1).copy()
    n = int(header["n_body"][0])
    have = (len(raw) - HEADER.itemsize) // FP.itemsize
    if max_particles is not None:
        n = min(n, int(max_particles), have)
    if have < n or (max_particles is None and len(raw) != HEADER.itemsize + n * FP.itemsize):
        raise ValueError("%s: header says %d particles, file holds %.2f" % (path, n, (len(raw) - 128) / 344.0))
    return header, np.frombuffer(raw, dtype=FP, count=n, offset=HEADER.itemsize).copy()


def write_binary(path, header, fp):
    header = np.ascontiguousarray(header, dtype=HEADER)
    fp = np.ascontiguousarray(fp, dtype=FP)
    with open(path, "wb") as f:
        f.write(header.tobytes())
        f.write(fp.tobytes())


# ------------------------------------------------------------------ ASCII snapshot
def _e(x):
    return "%20.15e" % x


def read_ascii(path):
    """(header[1], ASCII[n]) of a snapNNNNNN.dat / initial-condition file."""
    with open(path) as f:
        h = f.readline().split()
        header = np.zeros(1, dtype=HEADER)
        header["time"], header["n_body"], header["id_next"] = float(h[0]), int(h[1]), int(h[2])
        for k, name in enumerate(_E_ASCII):
            header["e_init"][name] = float(h[3 + k])
            header["e_now"][name] = float(h[8 + k])
        n = int(header["n_body"][0])
        p = np.zeros(n, dtype=ASCII)
        for i in range(n):
            t = f.readline().split()
            if len(t) != 12:
                raise ValueError("%s: particle line %d has %d fields" % (path, i, len(t)))
            p["id"][i] = int(t[0])
            p["mass"][i], p["r_planet"][i], p["f"][i] = float(t[1]), float(t[2]), float(t[3])
            p["pos"][i] = [float(x) for x in t[4:7]]
            p["vel"][i] = [float(x) for x in t[7:10]]
            p["n_neighbor"][i], p["flag"][i] = int(t[10]), int(t[11])
    return header, p


def write_ascii(path, header, p):
    """Byte-for-byte what FileHeader::writeAscii + FPGrav::writeAscii print."""
    h = header[0] if getattr(header, "shape", ()) else header
    with open(path, "w") as f:
        f.write("%g\t%d\t%d\t" % (h["time"], h["n_body"], h["id_next"]) +
                "\t".join(_e(h["e_init"][k]) for k in _E_ASCII) + "\t" +
                "\t".join(_e(h["e_now"][k]) for k in _E_ASCII) + "\n")
        for r in p:
            f.write("%d\t" % r["id"] + "\t".join(_e(x) for x in (r["mass"], r["r_planet"], r["f"], *r["pos"], *r["vel"])) +
                    "\t%d\t%d\n" % (r["n_neighbor"], r["flag"]))


def fp_to_ascii(fp):
    p = np.zeros(len(fp), dtype=ASCII)
    for k in ("id", "mass", "r_planet", "f", "pos", "vel"):
        p[k] = fp[k]
    p["n_neighbor"] = fp["neighbor"]["number"]
    return p


# ------------------------------------------------------------------ to and from the resident state
def fp_to_epj(fp):
    """EPJGrav[n] for gplum_b200.state.upload: the leading 112 bytes of every record, with id_local = slot."""
    raw = np.ascontiguousarray(fp, dtype=FP).view(np.uint8).reshape(len(fp), FP.itemsize)
    epj = np.ascontiguousarray(raw[:, :S.EPJ.itemsize]).view(S.EPJ).reshape(len(fp)).copy()
    epj["id_local"] = np.arange(len(fp))
    return epj


def epj_into_fp(epj, fp, time=None, dt=None):
    """Write a downloaded state back into restart records (pos, vel, acc_d; optionally time, dt)."""
    fp = fp.copy()
    for k in ("pos", "vel", "acc_d", "r_out", "r_search", "mass"):
        fp[k] = epj[k]
    fp["r_out_inv"] = 1.0 / epj["r_out"]
    if time is not None:
        fp["time"] = time
    if dt is not None:
        fp["dt"] = dt
    return fp


if __name__ == "__main__":
    import sys
    for path in sys.argv[1:]:
        with open(path, "rb") as f:
            head = f.read(64)
        if all(32 <= b < 127 or b in (9, 10, 13) for b in head):
            h, p = read_ascii(path)
        else:
            h, p = read_binary(path)
        print("%s: t=%g n_body=%d id_next=%d etot=%.15e  mass=[%.3e, %.3e]" % (
            path, h["time"][0], h["n_body"][0], h["id_next"][0], h["e_now"]["etot"][0], p["mass"].min(), p["mass"].max()))
