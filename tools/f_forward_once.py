"""ncu target: two forward passes of the surrogate (5 GEMM launches each) at batch 65 536."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from core.models.forward_model import ForwardModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
torch.manual_seed(0)
F = ForwardModel(4, 250, 8).cuda().eval()
p = torch.rand(B, 4, device="cuda") * 2 - 1
with torch.no_grad():
    for _ in range(2):
        s, m = F(p)
    torch.cuda.synchronize()
print("ok", float(s.float().abs().mean()))
