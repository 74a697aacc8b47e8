"""Absolute timeline (cycles since the first stamp) of both epilogue groups of CTA 0 for one F hidden layer."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from core.models.forward_model import ForwardModel
from pigan_b200 import native
B = 65536
torch.manual_seed(0)
F = ForwardModel(4, 250, 8).cuda().eval()
p = torch.rand(B, 4, device="cuda") * 2 - 1
with torch.no_grad():
    F(p); torch.cuda.synchronize()
    tr = torch.zeros(4, 64, 5, dtype=torch.int64, device="cuda")
    native.lib.pigan_engine_trace_layernorm(tr.data_ptr())
    F(p); torch.cuda.synchronize()
    native.lib.pigan_engine_trace_layernorm(None)
    li = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    t = tr[li].cpu()
    t0 = int(t[:, 0][t[:, 0] > 0].min())
    for row in range(16):
        if t[row, 0] == 0: continue
        g = row & 1
        a = [int(x) - t0 for x in t[row, :4]]
        print(f"unit {row >> 1} group {g}: start {a[0]:7d}  pass1 done {a[1]:7d}  exchanged {a[2]:7d}  pass2 done {a[3]:7d}")
