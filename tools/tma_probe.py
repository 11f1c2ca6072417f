import sys, numpy as np
sys.path.insert(0, "/root/repo")
from shll_sve_cfd_b200 import capi, programs
pb = programs.BASE_SHLL_2D.resized(64, 64)
u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
with programs.make_solver(pb) as s:
    print(s.variant)
    s.upload_u(u0)
    s.run(2)
    u = s.download_u()
    print("ok", float(u.sum()))
