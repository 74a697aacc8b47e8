"""Summarise an ncu launch list (gpu__time_duration.sum CSV): the launches between two occurrences of a marker
kernel.  usage: launch_summary.py FILE [WHICH [MARKER]] — MARKER 'center_vec' (default) opens a PI-GAN train step;
'end:NAME' takes the launches after occurrence WHICH of NAME up to and including occurrence WHICH+1 (e.g.
end:f_train_losses = one surrogate-training step of tools/one_step.py)."""
import csv, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; data = rows[hi + 1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
seq = []
for r in data:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    seq.append((r[ki], v))
marker = sys.argv[3] if len(sys.argv) > 3 else 'center_vec'
if marker.startswith('end:'):
    idx = [i for i, (k, _) in enumerate(seq) if marker[4:] in k]
    s, e = idx[which] + 1, idx[which + 1] + 1
else:
    idx = [i for i, (k, _) in enumerate(seq) if marker in k]
    s, e = idx[which], idx[which + 1] if which + 1 < len(idx) else len(seq)
tot = 0
for k, v in seq[s:e]:
    name = re.sub(r'pigan::(<unnamed>::)?', '', k)
    name = re.sub(r'\(.*', '', name)
    name = name.replace('GemmCfg', 'Cfg')
    print(f"{v:9.1f} us  {name[:110]}")
    tot += v
print(f"total {tot:.1f} us over {e - s} launches")
