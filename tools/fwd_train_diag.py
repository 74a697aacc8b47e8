"""Per-tensor gradient error of the surrogate-training step against the oracle (GPU box only)."""
import copy, os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_fwd_train as T
from oracle import models as O
from pigan_b200 import native
for B in [int(x) for x in (sys.argv[1:] or ["64", "1024", "4096"])]:
    for p in (0.2, 0.0):
        tr, F, f_sd, (pn, spec, mn) = T._setup(B, dropout_p=p)
        dump = torch.zeros(sum(T.HID) * B, dtype=torch.uint8, device="cuda")
        a = native.PiganFwdTrainArgs()
        png, sg, mng = pn.cuda(), spec.cuda(), mn.cuda()
        a.params_norm, a.spectrum, a.metrics_norm = png.data_ptr(), sg.data_ptr(), mng.data_ptr()
        a.batch = a.global_batch = B; a.first_row = 0
        a.f_params = tr.fs.params.tensor().data_ptr()
        a.f_grads, a.f_exp_avg, a.f_exp_avg_sq = tr.grads.data_ptr(), tr.m.data_ptr(), tr.v.data_ptr()
        a.lr, a.step, a.beta1, a.beta2, a.eps, a.max_norm = 1e-3, 1, 0.9, 0.999, 1e-8, 1.0
        a.dropout_p, a.dropout_seed = p, 1234
        a.losses, a.loss_sums, a.mask_dump = tr.losses.data_ptr(), tr.loss_sums.data_ptr(), dump.data_ptr()
        ws, nb, st = tr.workspace.data_ptr(), tr.workspace.numel(), native.current_stream()
        native.check(native.lib.pigan_fwd_train_step_phase(tr.engine.handle, C.byref(a), 0, ws, nb, st))
        torch.cuda.synchronize()
        masks = T._masks(dump, B)
        if p == 0.0:
            masks = [m * 0.8 for m in masks]      # the oracle divides by 0.8: 0.8 / 0.8 = identity
        ref_sd = copy.deepcopy(f_sd)
        ref, rg = O.pretrain_step(ref_sd, O.Adam(T._names(), betas=(0.9, 0.999)), pn, spec, mn, 1e-3, masks)
        views = dict(zip(T._names(), tr.fs.params.views_like(tr.grads)))
        flat = torch.cat([rg[n].reshape(-1) for n in T._names()])
        print(f"B={B} p={p} total grad rel {T.rel(tr.grads, flat):.2e}  |g|={float(flat.norm()):.3e}")
        print("   " + " ".join(f"{n.replace('model.', '')}:{T.rel(views[n], rg[n]):.1e}" for n in T._names()))
