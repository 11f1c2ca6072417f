"""CPU: run the compiled reference programs from oracle/_ref live (when present) and compare the oracle with them,
including step-capped prefixes of bigger grids that are too large to keep as fixtures."""
import numpy as np
import pytest

from conftest import ORACLE_IC, bits, oracle_cfg_for
from shll_sve_cfd_b200 import capi, programs

LIVE = [
    # (binary, problem, step cap)
    ("ref_1d_o1_8192", programs.BASE_SHLL.resized(8192), 500),
    ("ref_2d_o1_256", programs.BASE_SHLL_2D, 12),
    ("ref_2d_o2_128", programs.SECOND_ORDER_2D.resized(128), 40),
    ("ref_2d_o2_256", programs.SECOND_ORDER_2D, 6),
    ("ref_omp_o2_128", programs.BASE_OMP_2D.resized(128), 25),
]


@pytest.mark.parametrize("exe,pb,cap", LIVE, ids=[x[0] for x in LIVE])
def test_oracle_vs_live_reference(exe, pb, cap, oracle):
    O = oracle
    if not O.ref_available(exe):
        pytest.skip(f"oracle/_ref/{exe} not built (needs /root/reference)")
    r = O.run_ref(exe, pb.ncomp, pb.ncells, step_cap=cap, threads=2 if "omp" in exe else None)
    cfg = oracle_cfg_for(O, pb, nthreads=2)
    u = O.run(cfg, O.cons_from_prim(cfg, O.init_prim(cfg, ORACLE_IC[pb.ic])), cap)
    assert np.array_equal(bits(u), bits(r["u"]))


def test_derived_1d_second_order_is_the_y_uniform_slice(oracle):
    """the derived 1D 2nd-order program == any j-column of the y-uniform 2D 2nd-order run."""
    O = oracle
    exe = "ref_1d_o2_slice_4096"
    if not O.ref_available(exe):
        pytest.skip("oracle/_ref not built")
    n, cap = 4096, 300
    r = O.run_ref(exe, 4, n * 4, step_cap=cap)
    ru = r["u"].reshape(4, n, 4)
    assert not ru[2].any()
    pb = programs.SECOND_ORDER_1D.resized(n)
    cfg = oracle_cfg_for(O, pb)
    u = O.run(cfg, O.cons_from_prim(cfg, O.init_prim(cfg, O.IC_SOD_1D)), cap)
    for k1, k2 in enumerate((0, 1, 3)):
        for j in range(4):
            assert np.array_equal(bits(u[k1]), bits(ru[k2][:, j]))
