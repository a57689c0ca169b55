"""Config 1 end to end: the UNMODIFIED GPLUM program (INIT3000, perfect-merger collisions) built
(a) as shipped and (b) against libgplum_b200 through the drop-in headers -- per-call functor form
(include/pikg/*.hpp, -DUSE_PIKG), batched multi-walk form (include/gravity_kernel_b200.hpp) and the whole stage on
the GPU (include/gravity_tree_b200.hpp: Tree_t = gplum_b200::TreeB200 -- tree, lists, force pass, changeover
correction and neighbour lists from the device; FDPS's tree is never built).
The energy-error history, cluster statistics and mean neighbour count of the runs must agree
(BASELINE.json north_star: "energy-error histories of a full run agreeing with the reference").
Binaries are built in the build container by `make -C oracle ref` and travel in oracle/_ref/."""
import numpy as np
import pytest

import gplum_run as G

pytestmark = pytest.mark.gpu
need = pytest.mark.skipif(not (G.have("gplum_ref.out") and G.have("gplum_b200_mw.out")),
                          reason="oracle/_ref whole-program builds not present")


@need
@pytest.mark.parametrize("binary", ["gplum_b200_mw.out", "gplum_b200_functor.out", "gplum_b200_tree.out"])
def test_energy_history_agrees_with_reference_program(binary, tmp_path):
    if binary == "gplum_b200_tree.out" and not G.have(binary):
        pytest.skip("oracle/_ref/gplum_b200_tree.out not built")
    t_end = "2^-4" if binary.endswith("functor.out") else "2^-2"      # 4 / 16 tree steps
    ref, _ = G.run("gplum_ref.out", str(tmp_path / "ref"), t_end=t_end)
    # GPLUM_B200_FLAGS=1: the as-shipped quadrupole trace, i.e. the arithmetic of the binary we compare with
    got, out = G.run(binary, str(tmp_path / "b200"), t_end=t_end, env_extra={"GPLUM_B200_FLAGS": "1"})
    assert got.shape == ref.shape and len(ref) >= 5
    assert (got[:, 0] == ref[:, 0]).all() and (got[:, 1] == ref[:, 1]).all()          # time, n_tot
    assert np.abs(got[:, 2] / ref[:, 2] - 1).max() < 1e-10                              # etot
    assert np.abs(got[:, 3] - ref[:, 3]).max() < 5e-12, (got[:, 3], ref[:, 3])          # energy error history
    assert (got[:, 4:7] == ref[:, 4:7]).all()          # largest cluster, #clusters, #isolated particles
    assert np.allclose(got[:, 9], ref[:, 9], rtol=0, atol=1e-12)                        # mean neighbour number
