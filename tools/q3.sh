for acc in 0 1; do
echo "== SHLL_ACC=$acc"
SHLL_ACC=$acc python - <<'PY'
import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, json
from conftest import load_golden, problem_from_manifest
from shll_sve_cfd_b200 import capi, programs
man = json.load(open("tests/golden/MANIFEST.json"))
for case in ["1d_o1_256", "1d_o1_1024", "2d_o1_64", "2d_o1_96x160", "2d_o1_256", "2d_o2_64", "2d_o2_96x160", "1d_o2_slice_1024", "omp_o2_64"]:
    pb = problem_from_manifest(man[case])
    if case.startswith("1d_o2_slice"):
        pb = pb.resized(man[case]["nx"])
    try:
        gu, gp, gsteps = load_golden(case)
        r = programs.run_program(pb, capi.MODE_FAST)
        err = np.abs(r["p"].astype(np.float64) - gp.astype(np.float64).reshape(r["p"].shape))
        rel = err / (1 + np.abs(gp.astype(np.float64).reshape(r["p"].shape)))
        print(f"{case:18s} {r.get('variant','')[:40]:40s} max abs {err.max():.2e}  max err/(1+|ref|) {rel.max():.2e}  per-field {[f'{x:.1e}' for x in rel.max(axis=1)]}")
    except Exception as e:
        print(case, "ERR", repr(e)[:200])
PY
done
