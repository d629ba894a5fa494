# classification pre-pass, second version (one poll / one write of the verdict per block): parity, A/B, full bench
T=r06h
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_block_classified or benchmark_shapes or retry_policy or golden" 2>&1 | tail -3
timeout 60 python - <<'P'
import sys, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
for name,B in (('S0',400000),('S2',400000),('C3',100000)):
    b=workloads.make(name,B)
    t=[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in b.fields()]
    for label,fl in (('classified',0),('dense (NO_HOST_PACK)',32)):
        out=mgqp_b200.solve_quadprog_batched(*t, flags=fl); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): mgqp_b200.solve_quadprog_batched(*t, flags=fl, out=out)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        print(name,label,'%.3f ms -> %.4g QP/s'%(ms,B/ms*1e3), flush=True)
P
timeout 150 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
