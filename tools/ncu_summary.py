"""Compact per-kernel table from an ncu report (--set full): duration, DRAM bytes, DRAM %, tensor-pipe %, registers.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.csv"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration_us"), ("dram__bytes_read.sum", "dram_read_MB"),
        ("dram__bytes_write.sum", "dram_write_MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"), ("launch__grid_size", "grid"),
        ("sm__inst_executed.sum", "warp_insts"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__inst_executed_pipe_alu.sum", "alu_insts"), ("sm__inst_executed_pipe_fma.sum", "fma_insts"),
        ("smsp__inst_executed_pipe_xu.sum", "xu_insts")]
idx = [(hdr.index(k) if k in hdr else -1, n) for k, n in want]
units = rows[1]
w = csv.writer(sys.stdout)
w.writerow([n for _, n in idx])
for r in rows[2:]:
    out = []
    for i, n in idx:
        v = r[i] if i >= 0 else ""
        if n == "kernel":
            v = re.sub(r"pigan::(<unnamed>::)?", "", v)
            v = re.sub(r"\(.*", "", v).replace("GemmCfg", "Cfg").replace("(int)", "").replace("(bool)", "")
        elif n in ("dram_read_MB", "dram_write_MB") and i >= 0:
            u = units[i]
            f = float(v.replace(",", "")) if v else 0.0
            f = f / 1e6 if u == "byte" else (f / 1e3 if u == "Kbyte" else (f if u == "Mbyte" else f * 1e3))
            v = f"{f:.2f}"
        elif n == "duration_us" and i >= 0:
            u = units[i]
            f = float(v.replace(",", ""))
            f = f / 1e3 if u in ("ns", "nsecond") else (f * 1e3 if u in ("ms", "msecond") else f)
            v = f"{f:.2f}"
        out.append(v)
    w.writerow(out)
