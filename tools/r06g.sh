# where the classification path's 0.35 ms goes: per-launch durations (ncu metrics pass)
cat > /tmp/p.py <<'P'
import sys; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
for name,B in (('S2',400000),('C3',100000)):
    b=workloads.make(name,B)
    t=[torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in b.fields()]
    out=mgqp_b200.solve_quadprog_batched(*t)
    for _ in range(2): mgqp_b200.solve_quadprog_batched(*t, out=out)
    torch.cuda.synchronize()
P
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -c 60 --csv --log-file gpurun_out/r06g_launches.csv python /tmp/p.py > gpurun_out/r06g.log 2>&1
tail -2 gpurun_out/r06g.log
