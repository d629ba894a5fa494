# group kernel with the compressed CI block kept compressed in shared memory (15 warps per SM on the shipped shapes)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'P'
import sys, time; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
for name,B in (('S0',400000),('S2',400000)):
    b=workloads.make(name,B)
    val,row=mgqp_b200.pack_bound_columns(b.CI)
    t=[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in b.fields()]
    v=torch.from_numpy(val).to(dev); r=torch.from_numpy(row).to(dev)
    for label,kw,ci in (('dense',{},t[4]),('compressed',dict(CI_val=v,CI_row=r),None)):
        out=mgqp_b200.solve_quadprog_batched(t[0],t[1],t[2],t[3],ci,t[5],**kw); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): mgqp_b200.solve_quadprog_batched(t[0],t[1],t[2],t[3],ci,t[5],out=out,**kw)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        print(name,label,'%.3f ms -> %.4g QP/s'%(ms,B/ms*1e3), flush=True)
import bench
rec=bench.cascade_record(mgqp_b200, dev) if hasattr(bench,'cascade_record') else None
if rec: print('cascade', {k:(v['value'], v['e2e']['value'], v['parity_check']['tracking_bit_identical']) for k,v in rec.items()})
P
