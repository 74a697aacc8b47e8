"""Where does the end-to-end loop lose time?  Variants of the bench loop at N=1."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from core.models.generator import Generator
from core.models.discriminator import Discriminator
from core.models.forward_model import ForwardModel
from pigan_b200 import synthetic
from pigan_b200.trainer import NativeTrainer
B, K = 65536, 30
dev = torch.device("cuda")
torch.manual_seed(42)
G = Generator(250, 4); D = Discriminator(250, 4); F = ForwardModel(4, 250, 8); F.eval()
tr = NativeTrainer(G, D, F, dev, max_batch=B)
sp, pr, pn, mn = synthetic.make_batch(B, 250, seed=1, device=dev)
center = sp[:512].mean(0).contiguous()
op = NativeTrainer.prepare_operand(sp, pr, center)
hop, hmn = op.cpu().pin_memory(), mn.cpu().pin_memory()
dop, dmn = [torch.empty_like(op) for _ in range(2)], [torch.empty_like(mn) for _ in range(2)]
cs = torch.cuda.Stream(); main = torch.cuda.current_stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lh = torch.empty(K, 9).pin_memory()
def timeit(name, body):
    for _ in range(3): body(0)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    for i in range(K): body(i)
    e1.record(); th = time.perf_counter() - t0; torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/K:.3f} ms/step (host loop {th/K*1e3:.3f} ms/step)")
timeit("device only (prepared)", lambda i: tr.step_prepared(op, center, mn, 2e-4, 2e-4))
timeit("device only + D2H losses", lambda i: lh[i].copy_(tr.step_prepared(op, center, mn, 2e-4, 2e-4), non_blocking=True))
def copy_only(i):
    with torch.cuda.stream(cs):
        dop[i % 2].copy_(hop, non_blocking=True); dmn[i % 2].copy_(hmn, non_blocking=True)
timeit("H2D only (35.7 MB, copy stream)", lambda i: (copy_only(i), main.wait_stream(cs)))
def both(i):
    copy_only(i)
    tr.step_prepared(op, center, mn, 2e-4, 2e-4)
timeit("step + unsynchronised concurrent H2D", lambda i: (both(i), main.wait_stream(cs)) if i == K - 1 else both(i))
