#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/wp_dbg.py <<'PY'
import os, sys, subprocess, tempfile
sys.path.insert(0, "tests")
import gplum_run as G
n, group, steps_exp = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
d = tempfile.mkdtemp(prefix="gplum_dbg_")
p = dict(G.PARAMS, makeInit="1", n_init=str(n), n_group_limit=str(group), t_end=steps_exp, dt_snap="1", dt_snap_tmp="1")
with open(os.path.join(d, "param.dat"), "w") as f:
    for k, v in p.items(): f.write("%-16s= %s\n" % (k, v))
import resource
env = dict(os.environ, OMP_NUM_THREADS="16", GPLUM_B200_FLAGS="1", OMP_STACKSIZE="1G")
def big_stack():
    resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
r = subprocess.run([os.path.join(G.REF_DIR, sys.argv[1]), "-p", "param.dat"], cwd=d, env=env, capture_output=True, text=True, timeout=900, preexec_fn=big_stack)
out = [l for l in r.stdout.splitlines() if ("Time:" in l or "EnergyError" in l or "Wall Time" in l)]
print(sys.argv[1], n, group, "rc", r.returncode); print("\n".join(out[-12:])); print(r.stderr[-400:])
PY
for b in gplum_b200_tree.out; do timeout 900 python /tmp/wp_dbg.py $b 100000 64 2^-4; done > gpurun_out/wp_debug.log 2>&1
cat gpurun_out/wp_debug.log | cut -c1-220
