"""Does the force pass need the DSL's Newton step after MUFU.RSQ?  (kernels.cuh: GB_NEWTON)

  run  <tag>   one N=1e6 pass with the library GPLUM_B200_LIB names (default: the shipped one): time per pass,
               forces to gpurun_out/newton_<tag>.npy
  compare a b  both results and the oracle's (the reference's FP32 arithmetic, rsqrt + Newton step) against the same
               sums evaluated in FP64 (numpy) on every 100th walk; CPU only, reads the two .npy files

Test infrastructure (reads oracle/): never imported by the product."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gplum_b200 import disk, tree


def workload():
    d = disk.make_disk(1000000)
    ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
    w, _ = tree.build_walks(d["pos"], d["mass"], ro, rs, n_group_limit=512)
    return w


def truth(w, k):
    """FP64 sums of walk k: src/gravity_kernel_epep.pikg:53-97, gravity_kernel_epsp.pikg:47-100, eps2 = 0"""
    i0, ni = w.epi_off[k], w.ni[k]
    xi = w.epi["pos"][i0:i0 + ni]; roi = w.epi["r_out"][i0:i0 + ni]
    ej = w.epj_all[w.adr_epj[w.epj_disp[k]:w.epj_disp[k] + w.n_epj[k]]]
    d = ej["pos"][None, :, :] - xi[:, None, :]
    r2 = (d * d).sum(-1)
    ro = np.maximum(roi[:, None], ej["r_out"][None, :])
    y = 1.0 / np.sqrt(np.maximum(r2, ro * ro))
    acc = ((ej["mass"][None, :] * y ** 3)[:, :, None] * d).sum(1)
    phi = -(ej["mass"][None, :] * y).sum(1)
    sj = w.spj_all[w.adr_spj[w.spj_disp[k]:w.spj_disp[k] + w.n_spj[k]]]
    if len(sj):
        d = sj["pos"][None, :, :] - xi[:, None, :]
        r2 = (d * d).sum(-1)
        y = 1.0 / np.sqrt(r2)
        q = sj["quad"]                                   # xx yy zz xy xz yz
        tr = q[:, 0] + q[:, 1] + q[:, 2]
        Q = np.zeros((len(sj), 3, 3))
        Q[:, 0, 0], Q[:, 1, 1], Q[:, 2, 2] = 3 * q[:, 0] - tr, 3 * q[:, 1] - tr, 3 * q[:, 2] - tr
        Q[:, 0, 1] = Q[:, 1, 0] = 3 * q[:, 3]; Q[:, 0, 2] = Q[:, 2, 0] = 3 * q[:, 4]; Q[:, 1, 2] = Q[:, 2, 1] = 3 * q[:, 5]
        qr = np.einsum("jab,ijb->ija", Q, d)
        rqr = (qr * d).sum(-1)
        m = sj["mass"][None, :]
        meff = m + 0.5 * rqr * y ** 4
        meff3 = (m + 2.5 * rqr * y ** 4) * y ** 3
        phi = phi - (meff * y).sum(1)
        acc = acc + (meff3[:, :, None] * d - qr * (y ** 5)[:, :, None]).sum(1)
    return acc, phi


def main():
    mode = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    w = workload()
    if mode == "run":
        from gplum_b200 import functors as F
        F.init(0); F.set_params(0.0, True, 0)
        F.walks_upload(w)
        F.walks_run(repack=True)
        f = F.walks_download(len(w.epi))
        ms = F.walks_time(30, repack=False)
        np.save(os.path.join(ROOT, "gpurun_out", "newton_%s.npy" % sys.argv[2]), f)
        print("%s: %.4f ms per pass (force kernel alone, 30 passes)" % (sys.argv[2], ms))
        return
    import oracle_api as O
    res = {t: np.load(os.path.join(ROOT, "gpurun_out", "newton_%s.npy" % t)) for t in sys.argv[2:4]}
    res["oracle (C restatement of the DSL, scalar, rsqrt + Newton as written)"] = O.calc_walks(w, 0.0)[0]
    sel = range(0, w.n_walk, 100)
    T = [truth(w, k) for k in sel]
    print("relative error against the FP64 sums, %d walks (%d particles): |d acc| / |acc|, |d phi| / |phi|"
          % (len(T), sum(len(t[1]) for t in T)))
    for name, f in res.items():
        ea, ep = [], []
        for k, (acc, phi) in zip(sel, T):
            i0, ni = w.epi_off[k], w.ni[k]
            ea.append(np.linalg.norm(f["acc"][i0:i0 + ni] - acc, axis=1) / np.linalg.norm(acc, axis=1))
            ep.append(np.abs(f["phi"][i0:i0 + ni] - phi) / np.abs(phi))
        ea, ep = np.concatenate(ea), np.concatenate(ep)
        print("  %-72s acc: median %.2e  99%% %.2e  max %.2e   phi: median %.2e  max %.2e"
              % (name, np.median(ea), np.percentile(ea, 99), ea.max(), np.median(ep), ep.max()))
    a, b = (res[t] for t in sys.argv[2:4])
    nb = all(np.array_equal(a[k], b[k]) for k in ("number", "rank", "id_max", "id_min"))
    print("neighbour candidates identical between %s and %s: %s" % (sys.argv[2], sys.argv[3], nb))


if __name__ == "__main__":
    main()
