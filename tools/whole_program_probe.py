"""The unmodified GPLUM program, as shipped and with Tree_t = gplum_b200::TreeB200 (include/gravity_tree_b200.hpp), on a
disk the program generates itself (makeInit = 1): wall time of the run and of its soft part from the program's own
timer (src/time.h: "Wall Time .. Soft .. Hard"), 64 tree steps (t = 0 .. 1).
Usage: python tools/whole_program_probe.py <n>
Binaries (oracle/_ref, built by `make -C oracle ref`): gplum_ref_simd.out = the reference's AVX2 flag set (-O3 -mavx2
-mfma -ffast-math), gplum_b200_tree.out = as shipped (-O2) + the two-line TreeB200 patch.  Test infrastructure."""
import os, re, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gplum_run as G

n = int(sys.argv[1])
threads = os.cpu_count() or 1
runs = [("gplum_ref_simd.out", 64, None), ("gplum_ref_simd.out", 512, None)]
if os.environ.get("WP_REF_ONLY") != "1":
    runs += [("gplum_b200_tree.out", 64, {"GPLUM_B200_FLAGS": "1"}), ("gplum_b200_tree.out", 512, {"GPLUM_B200_FLAGS": "1"})]
for binary, group, env in runs:
    if not G.have(binary):
        print(binary, "not built"); continue
    try:
        e, out = G.run_generated(binary, tempfile.mkdtemp(prefix="gplum_wp_"), n, t_end="1", dt_snap="1", threads=threads, env_extra=env,
                                 n_group_limit=group, timeout=1500)
    except Exception as ex:                                  # noqa: BLE001
        print(binary, group, "FAILED", str(ex)[-600:]); continue
    m = re.findall(r"Wall Time: ([0-9.eE+-]+)\s+Soft: ([0-9.eE+-]+)\s+Hard: ([0-9.eE+-]+)", out)
    tot, soft, hard = (float(x) for x in m[-1])
    print("%-20s n_group_limit=%-4d N=%d, 64 tree steps, %d host threads: wall %.3f s, Soft %.3f s (%.1f ms per step), Hard %.3f s; "
          "particles left %d, energy error at t=1 %.4e" % (binary, group, n, threads, tot, soft, soft / 64 * 1e3, hard, int(e[-1, 1]), e[-1, 3]))
