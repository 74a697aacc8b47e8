import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pigan_b200 import native_test as native
M, n, k = 65536, 512, 256
a = torch.randn(M, k, device="cuda").half(); b = torch.randn(n, k, device="cuda").half()
bias = torch.randn(n, device="cuda"); out = torch.empty(M, n, device="cuda", dtype=torch.float16)
st = native.current_stream()
for _ in range(3):
    native.check(native.lib.pigan_debug_linear(a.data_ptr(), b.data_ptr(), bias.data_ptr(), out.data_ptr(), None, None, None, M, n, k, st))
torch.cuda.synchronize()
print("ok")
