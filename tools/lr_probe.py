"""One cf_lowrank_project + reconstruct call per rank (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from compactfusion_b200.compress_lowrank import lowrank_project, lowrank_reconstruct
dev = torch.device("cuda:0")
n, c = 4608, 3072
x = torch.randn(n, c, device=dev).half()
base = (x.float() + 0.3 * torch.randn(n, c, device=dev)).half()
for r in [int(a) for a in sys.argv[1:]] or [32]:
    u, v, _ = lowrank_project(x, base, r, 2)
    lowrank_reconstruct(u, v, base)
torch.cuda.synchronize()
print("ok")
