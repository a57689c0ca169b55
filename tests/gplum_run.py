"""Run a whole-program GPLUM build (oracle/_ref/*.out) on config 1 (INIT3000, perfect-merger
collisions) in a scratch directory and return its energy.dat.  The input file is re-created
from tests/golden/init3000_particles.npz in the reference's ASCII format
(src/particle.h:844-858, header src/energy.h:109-118); the parameter file carries the keys of
sample/parameter.dat with makeInit=0."""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(os.path.dirname(HERE), "oracle", "_ref")

PARAMS = dict(
    seed="0", init_file="INIT3000.dat", Header="1", output_dir="TEST", Restart="0", makeInit="0",
    n_init="3000", m_init="2.e22CGS", p="1.5", f_dust="0.71", eta_ice="30./7.1", a_in="0.9", a_out="1.1",
    a_ice="2.0", ecc_hill="2.0", inc_hill="1.0", alpha_gas="11./4.", beta_gas="0.5", f_gas="0.71", tau_gas="0.",
    C_d="1.", mu="2.34", coef_ema="0.3", reset_step="1024", theta="0.5", n_leaf_limit="8", n_group_limit="64",
    n_smp_ave="100", t_end="1", dt_tree="2^-6", dt_snap="1", dt_snap_tmp="1", dt_min="2^-30", eta="0.02",
    eta_sun="0.02", eta_0="0.002", eta_sun0="0.002", alpha="1.", m_sun="1.", dens="2.CGS", eps="0.", eps_sun="0.",
    R_cut0="3.0", R_cut1="8.0", R_search0="1.1", R_search1="6.0", R_search2="1.1", R_search3="2.0", R_merge="0.2",
    gamma="0.5", r_cut_max="0.", r_cut_min="0.", p_cut="0.", r_max="20.", r_min="0.1", f="1.", m_min="2.e22/10.CGS",
    a_frag="0.0", N_frag="10", dens_imp="1.CGS", c_s="1.8", mu_="1./3.", eta_="-3./2.", eps_n="1.", eps_t="1.")


def have(binary):
    return os.path.exists(os.path.join(REF_DIR, binary))


def write_inputs(d, t_end, dt_snap):
    z = np.load(os.path.join(HERE, "golden", "init3000_particles.npz"))
    h = z["header"]
    with open(os.path.join(d, "INIT3000.dat"), "w") as f:
        f.write("%g\t%d\t%d\t" % (h[0], int(h[1]), int(h[2])) + "\t".join("%20.15e" % x for x in h[3:]) + "\n")
        for i in range(len(z["id"])):
            row = [z["mass"][i], z["r_planet"][i], z["f"][i], *z["pos"][i], *z["vel"][i]]
            f.write("%d\t" % z["id"][i] + "\t".join("%20.15e" % x for x in row) + "\t0\t0\n")
    p = dict(PARAMS, t_end=t_end, dt_snap=dt_snap, dt_snap_tmp=dt_snap)
    with open(os.path.join(d, "param.dat"), "w") as f:
        for k, v in p.items():
            f.write("%-16s= %s\n" % (k, v))


def run(binary, workdir, t_end="2^-2", dt_snap="2^-6", threads=4, env_extra=None, timeout=900):
    os.makedirs(workdir, exist_ok=True)
    write_inputs(workdir, t_end, dt_snap)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    env.update(env_extra or {})
    r = subprocess.run([os.path.join(REF_DIR, binary), "-p", "param.dat"], cwd=workdir, env=env,
                       capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("%s failed (%d):\n%s\n%s" % (binary, r.returncode, r.stdout[-2000:], r.stderr[-2000:]))
    e = np.loadtxt(os.path.join(workdir, "TEST", "energy.dat"), ndmin=2)
    return e, r.stdout


def _big_stack():
    # FDPS keeps per-thread work arrays on the stack; at N >= 1e5 the default 8 MB overflows (the reference itself)
    import resource
    resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))


def run_generated(binary, workdir, n, t_end="2^-2", dt_snap="2^-4", threads=8, env_extra=None, timeout=900, n_group_limit=64):
    """The same program on a disk it generates itself (makeInit = 1, n particles, seed 0): energy.dat rows and stdout."""
    os.makedirs(workdir, exist_ok=True)
    p = dict(PARAMS, makeInit="1", n_init=str(n), n_group_limit=str(n_group_limit), t_end=t_end, dt_snap=dt_snap, dt_snap_tmp=dt_snap)
    with open(os.path.join(workdir, "param.dat"), "w") as f:
        for k, v in p.items():
            f.write("%-16s= %s\n" % (k, v))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_STACKSIZE="1G")
    env.update(env_extra or {})
    r = subprocess.run([os.path.join(REF_DIR, binary), "-p", "param.dat"], cwd=workdir, env=env,
                       capture_output=True, text=True, timeout=timeout, preexec_fn=_big_stack)
    if r.returncode != 0:
        raise RuntimeError("%s failed (%d):\n%s\n%s" % (binary, r.returncode, r.stdout[-2000:], r.stderr[-2000:]))
    e = np.loadtxt(os.path.join(workdir, "TEST", "energy.dat"), ndmin=2)
    return e, r.stdout
