import sys, numpy as np
sys.path.insert(0,'/root/repo')
from gplum_b200 import disk, tree, functors as F
F.init(0); F.set_params(0.0, True, 0)
n=1000000
d = disk.make_disk(n); ro, rs = disk.cutoff_radii(d["pos"], d["vel"], d["mass"])
for _ in range(3):
    sz = tree.build_walks_gpu(d["pos"], d["mass"], ro, rs, n_group_limit=512)
print(tree.gpu_build_times())
print(tree.gpu_build_stamps())
