"""torchrun --nproc-per-node N tools/dp_check.py : peer-memory exchange vs NCCL all-reduce, same data and weights."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from core.models.generator import Generator
from core.models.discriminator import Discriminator
from core.models.forward_model import ForwardModel
from pigan_b200 import synthetic
from pigan_b200.trainer import NativeTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
res = {}
for mode in ("nccl", "peer"):
    os.environ["PIGAN_DP_EXCHANGE"] = mode
    torch.manual_seed(42)
    G = Generator(250, 4); D = Discriminator(250, 4); F = ForwardModel(4, 250, 8); F.eval()
    tr = NativeTrainer(G, D, F, dev, max_batch=B)
    assert (tr.xchg is not None) == (mode == "peer"), (mode, tr.xchg)
    ls = []
    for s in range(4):
        sp, pr, pn, mn = synthetic.make_batch(B, 250, seed=100 * rank + s, device=dev)
        ls.append(tr.step(sp, pr, mn, 2e-4, 2e-4).clone())
    torch.cuda.synchronize()
    res[mode] = (torch.stack(ls).cpu(), tr.gs.params.tensor().clone().cpu(), tr.ds.params.tensor().clone().cpu(),
                 tr.gs.bn.tensor().clone().cpu())
    # replicas identical across ranks?
    chk = tr.gs.params.tensor().double().sum().reshape(1)
    allc = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    assert all(torch.equal(allc[0], c) for c in allc), f"{mode}: replicas diverged"
    # timing
    sp, pr, pn, mn = synthetic.make_batch(B, 250, seed=7 + rank, device=dev)
    for _ in range(3): tr.step(sp, pr, mn, 2e-4, 2e-4)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): tr.step(sp, pr, mn, 2e-4, 2e-4)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{mode}: {e0.elapsed_time(e1)/20:.3f} ms/step at B={B} x {world}")
    if mode == "peer":
        tr.exchange_events = []
        for _ in range(10): tr.step(sp, pr, mn, 2e-4, 2e-4)
        torch.cuda.synchronize()
        acc = {}
        for name, a, b in tr.exchange_events:
            acc.setdefault(name, []).append(a.elapsed_time(b) * 1e3)
        tr.exchange_events = None
        if rank == 0:
            print("  per-exchange us on rank 0 (kernel + wait for the slowest peer):",
                  {k: round(sum(v) / len(v), 1) for k, v in acc.items()})
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
if rank == 0:
    print("losses rel", rel(res["peer"][0], res["nccl"][0]), "G params", rel(res["peer"][1], res["nccl"][1]),
          "D params", rel(res["peer"][2], res["nccl"][2]), "BN", rel(res["peer"][3], res["nccl"][3]))
    assert rel(res["peer"][0], res["nccl"][0]) < 1e-5 and rel(res["peer"][1], res["nccl"][1]) < 1e-4
    print("dp_check ok")
dist.destroy_process_group()
