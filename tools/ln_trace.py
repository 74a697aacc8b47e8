"""Cycle stamps of the LayerNorm epilogue (CTA 0, group 0, thread 0) for the four F hidden layers."""
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
    F(p)
    torch.cuda.synchronize()
    import ctypes
    tr = torch.zeros(4, 64, 5, dtype=torch.int64, device="cuda")
    native.lib.pigan_engine_trace_layernorm(tr.data_ptr())
    F(p)
    torch.cuda.synchronize()
    native.lib.pigan_engine_trace_layernorm(None)
    for li, name in enumerate(["L2 256->512", "L3 512->1024 (cluster 2)", "L4 1024->512", "L5 512->256"]):
        t = tr[li].cpu()[0::2]        # group 0's units (group 1's are the odd rows)
        k = int((t[:, 0] > 0).sum())
        d = (t[:k, 1:4] - t[:k, 0:3]).float()
        gaps = (t[1:k, 0] - t[:k - 1, 3]).float()
        print(f"{name}: units {k}; pass1 {d[:,0].mean():.0f}  exchange {d[:,1].mean():.0f}  pass2 {d[:,2].mean():.0f} cycles;"
              f" wait for next accumulator {gaps.mean():.0f}; unit period {(t[k-1,0]-t[0,0]).item()/(k-1):.0f}")
