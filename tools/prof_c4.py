"""One launch of the CTA-per-QP kernel on device-generated C4 QPs (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mgqp_b200
from mgqp_b200 import workloads
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
t = workloads.make_c4_device(B, torch.device("cuda:0"))
torch.cuda.synchronize()
out = mgqp_b200.solve_quadprog_batched(*t, flags=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
torch.cuda.synchronize()
print("done", out.status.bincount().tolist())
