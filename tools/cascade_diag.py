import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import mgqp_b200
from mgqp_b200 import workloads
from oracle import cascade_oracle as co, oracle_py
for cfg, scale in ((workloads.WIDE_CONFIG, 0.01), (workloads.WIDE_CONFIG, 0.05), (workloads.SHIPPED_CONFIG, 1.0)):
    s = workloads.make_scenarios(20000, seed=11, config=cfg, scale=scale)
    t = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    r = mgqp_b200.solve_cascade_batched(t, cfg, taps=True); torch.cuda.synchronize()
    got = {k: (v.cpu().numpy() if v is not None else None) for k, v in r._asdict().items()}
    ref = co.cascade(oracle_py.restatement(), s, cfg)
    a0 = got["status0"] == ref["status0"]; a2 = got["status2"] == ref["status2"]
    sc = np.maximum(1.0, np.abs(ref["tracking"]).max(axis=1))
    d = np.abs(got["tracking"] - ref["tracking"]).max(axis=1) / sc
    print("scale", scale, "agree0", a0.mean(), "agree2", a2.mean(), "Zdiff max", np.abs(got["Z"] - ref["Z"]).max())
    for tol in (1e-5, 1e-4, 1e-3, 1e-2):
        print("   tol", tol, "close", (d <= tol).mean(), "close|agree", (d[a0 & a2] <= tol).mean())
    bad = np.nonzero((a0 & a2) & (d > 1e-3))[0][:3]
    for i in bad:
        print("   example", i, "st", ref["status0"][i], ref["status2"][i], "d", d[i], "|u| ref", np.abs(ref["tracking"][i]).max())
