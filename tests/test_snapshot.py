"""Snapshot / restart wire formats (gplum_b200/snapshot.py) against files the reference program
itself wrote: the committed fixture (first 96 particles of snap_tmp.dat and of the ASCII snapshot of the same
instant, from oracle/_ref/gplum_ref.out on config 1) and, where oracle/_ref exists, a fresh run and the struct
layout reported by the compiled reference.  CPU only.  Bar: byte equality."""
import os

import numpy as np
import pytest

import oracle_api as O
from gplum_b200 import snapshot as SN, structs as S

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "snapshot_ref.npz")


def _fixture(tmp_path):
    z = np.load(GOLD)
    b, a = tmp_path / "snap_tmp.dat", tmp_path / "snap.dat"
    b.write_bytes(z["binary"].tobytes()); a.write_bytes(z["ascii"].tobytes())
    return z, str(b), str(a)


def test_binary_records_and_ascii_lines_of_the_reference(tmp_path):
    z, b, a = _fixture(tmp_path)
    keep = int(z["keep"])
    h, fp = SN.read_binary(b, max_particles=keep)
    assert int(h["n_body"][0]) == int(z["n_body"]) == 3000 and len(fp) == keep
    assert h["time"][0] == 2.0 ** -4 and (fp["time"] == 2.0 ** -4).all()
    assert np.array_equal(fp["id_local"], np.arange(keep)) and (fp["myrank"] == 0).all()
    assert np.array_equal(fp["r_out_inv"], 1.0 / fp["r_out"]) and (fp["r_search"] > fp["r_out"]).all()
    assert (fp["inDomain"] == 1).all() and (fp["isDead"] == 0).all()
    r = np.sqrt((fp["pos"] ** 2).sum(1))
    assert (0.85 < r).all() and (r < 1.15).all()
    # the restart records re-emitted: byte equal
    out = tmp_path / "rt.bin"
    SN.write_binary(str(out), h, fp)
    assert out.read_bytes() == z["binary"].tobytes()
    # the ASCII snapshot of the same instant: parse (header says 3000, the fixture holds `keep` lines)
    lines = z["ascii"].tobytes().split(b"\n")
    hdr = lines[0].split()
    assert int(hdr[1]) == 3000 and float(hdr[0]) == 2.0 ** -4
    # ... and the binary records printed the way FPGrav::writeAscii prints them give the same bytes
    h2 = h.copy(); h2["n_body"] = keep
    txt = tmp_path / "rt.dat"
    SN.write_ascii(str(txt), h2, SN.fp_to_ascii(fp))
    got = txt.read_bytes().split(b"\n")
    assert got[1:keep + 1] == lines[1:keep + 1]
    assert got[0].split()[3:] == hdr[3:]                      # the ten energies
    ha, pa = SN.read_ascii(str(txt))
    assert np.array_equal(pa["id"], fp["id"]) and np.abs(pa["pos"] - fp["pos"]).max() < 1e-15
    again = tmp_path / "rt2.dat"
    SN.write_ascii(str(again), ha, pa)
    assert again.read_bytes() == txt.read_bytes()


def test_records_feed_the_resident_state(tmp_path):
    z, b, a = _fixture(tmp_path)
    h, fp = SN.read_binary(b, max_particles=int(z["keep"]))
    epj = SN.fp_to_epj(fp)
    assert epj.dtype == S.EPJ
    for k in ("pos", "vel", "mass", "r_out", "r_search", "id", "acc_d"):
        assert np.array_equal(epj[k], fp[k]), k
    epj2 = epj.copy(); epj2["pos"] += 0.5; epj2["vel"] *= 2.0
    fp2 = SN.epj_into_fp(epj2, fp, time=fp["time"] + 1.0)
    assert np.array_equal(fp2["pos"], fp["pos"] + 0.5) and np.array_equal(fp2["time"], fp["time"] + 1.0)
    assert np.array_equal(fp2["acc"], fp["acc"]) and np.array_equal(fp2["neighbor"], fp["neighbor"])


def test_truncated_or_inconsistent_files_are_rejected(tmp_path):
    z, b, a = _fixture(tmp_path)
    with pytest.raises(ValueError):
        SN.read_binary(b)                        # header says 3000 particles, the fixture holds 96
    short = tmp_path / "short.dat"
    short.write_bytes(z["binary"].tobytes()[:100])
    with pytest.raises(ValueError):
        SN.read_binary(str(short))
    bad = tmp_path / "bad.dat"
    bad.write_bytes(b"\n".join(z["ascii"].tobytes().split(b"\n")[:5]) + b"\n1 2 3\n")
    with pytest.raises((ValueError, IndexError)):
        SN.read_ascii(str(bad))


@pytest.mark.skipif(not O.have_ref("scalar"), reason="oracle/_ref not built")
def test_layout_equals_the_compiled_reference():
    import ctypes as C
    out = (C.c_int * 64)()
    n = O.ref("scalar").ref_snapshot_layout(out)
    v = list(out[:n])
    assert v[0] == SN.HEADER.itemsize and v[1] == SN.ENERGY.itemsize
    assert v[2:7] == [SN.HEADER.fields[k][1] for k in ("n_body", "id_next", "time", "e_init", "e_now")]
    assert v[7] == SN.FP.itemsize
    assert v[8:] == [SN.FP.fields[k][1] for k in SN.FP.names]
    assert [SN.FP.fields[k][1] for k in S.EPJ.names] == [S.EPJ.fields[k][1] for k in S.EPJ.names]   # FPGrav : EPJGrav


@pytest.mark.skipif(not os.path.exists(os.path.join(HERE, "..", "oracle", "_ref", "gplum_ref.out")), reason="reference binary not built")
def test_fresh_reference_run_round_trips(tmp_path):
    import gplum_run
    d = str(tmp_path / "run")
    gplum_run.run("gplum_ref.out", d, t_end="2^-5", dt_snap="2^-5", threads=2)
    t = os.path.join(d, "TEST")
    h, fp = SN.read_binary(os.path.join(t, "snap_tmp.dat"))
    ha, pa = SN.read_ascii(os.path.join(t, "snap000001.dat"))
    assert len(fp) == len(pa) == 3000 and h["time"][0] == ha["time"][0] == 2.0 ** -5
    out = str(tmp_path / "a.dat")
    SN.write_ascii(out, ha, pa)
    assert open(out, "rb").read() == open(os.path.join(t, "snap000001.dat"), "rb").read()
    SN.write_ascii(out, h, SN.fp_to_ascii(fp))
    assert open(out, "rb").read() == open(os.path.join(t, "snap000001.dat"), "rb").read()
    out = str(tmp_path / "b.dat")
    SN.write_binary(out, h, fp)
    assert open(out, "rb").read() == open(os.path.join(t, "snap_tmp.dat"), "rb").read()
