"""Generates tests/golden/cascade_*.npz: robot states and what oracle/cascade_oracle.py (assembly + cascade, QP solves
by the pinned solver oracle) answers for them.  Run from the repository root:  python tests/golden/make_cascade_golden.py

PARITY UNPINNED against the reference for this fixture (see oracle/cascade_oracle.py: the callers are Eigen float32
code and Eigen is not in the reference tree); it pins the restatement against itself across platforms / numpy versions
and gives the GPU tests inputs whose expected answers do not depend on a LAPACK being present."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mgqp_b200 import workloads  # noqa: E402
from oracle import cascade_oracle as co, oracle_py  # noqa: E402

for name, cfg, scale, B in (("shipped", workloads.SHIPPED_CONFIG, 1.0, 96), ("wide", workloads.WIDE_CONFIG, 0.01, 96)):
    s = workloads.make_scenarios(B, seed=97, config=cfg, scale=scale)
    r = co.cascade(oracle_py.restatement(), s, cfg)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"cascade_{name}.npz"), **s,
                        tracking=r["tracking"], torques=r["torques"], status0=r["status0"], status2=r["status2"],
                        CE0=r["qp0"]["CE"], ce0=r["qp0"]["ce0"], ci0=r["qp0"]["ci0"], Z=r["Z"], res0=r["res0"])
    print(name, "status0", np.unique(r["status0"], return_counts=True), "status2", np.unique(r["status2"], return_counts=True))
