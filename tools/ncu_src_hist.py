"""Per-opcode and per-stall summary of one kernel of an `ncu --set full --import-source on` report (CPU, needs ncu).

    python tools/ncu_src_hist.py REPORT.ncu-rep LAUNCH_INDEX
"""
import collections, csv, io, re, subprocess, sys

rep, idx = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:200])
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
data = data[:len(data) // 2] if len(data) > 2000 and data[0][ci["Source"]] == data[len(data) // 2][ci["Source"]] else data
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
op = collections.Counter(); samp = collections.Counter(); st = collections.Counter()
tot = tots = 0
for r in data:
    try:
        n = int(r[ci["Instructions Executed"]]); s = int(r[ci["# Samples"]])
    except (ValueError, IndexError):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
    o = ".".join((m.group(2) if m else "?").split(".")[:2])
    op[o] += n; samp[o] += s; tot += n; tots += s
    for h in stalls:
        if r[ci[h]]:
            st[h] += int(r[ci[h]])
print(f"warp instructions {tot}, samples {tots}")
print("stalls:", ", ".join(f"{k[6:]} {v}" for k, v in st.most_common(12)))
for o, n in op.most_common(28):
    print(f"  {o:26s} {n:10d} {100 * n / tot:5.1f}%   samples {samp[o]:6d} {100 * samp[o] / max(tots, 1):5.1f}%")
