"""L2 / shared-memory / L1 throughput columns of an ncu --set full report, per kernel launch.
usage: python tools/ncu_l2.py report.ncu-rep > profiles/xxx.csv"""
import csv, io, re, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
pick = [i for i, h in enumerate(hdr) if h in ("Kernel Name", "gpu__time_duration.sum", "launch__grid_size") or
        re.match(r"(lts__t_bytes\.sum$|lts__t_sectors\.sum$|lts__throughput\.avg\.pct|lts__t_sectors_srcunit_tex\.sum$|"
                 r"lts__t_sectors_srcunit_tex_op_read\.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|"
                 r"l1tex__throughput\.avg\.pct|l1tex__m_xbar2l1tex_read_bytes\.sum$|sm__pipe_tensor_cycles_active\.avg\.pct|"
                 r"smsp__cycles_elapsed\.avg$|sm__cycles_elapsed\.avg$|sm__cycles_elapsed\.max$|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|smsp__cycles_active\.avg$|"
                 r"lts__t_sector_hit_rate\.pct$|gpc__cycles_elapsed\.max$|sm__cycles_active\.avg$)", h)]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in pick])
w.writerow([units[i] for i in pick])
for r in rows[2:]:
    out = []
    for i in pick:
        v = r[i]
        if hdr[i] == "Kernel Name":
            v = re.sub(r"pigan::(<unnamed>::)?", "", v)
            v = re.sub(r"\(.*", "", v).replace("GemmCfg", "Cfg").replace("(int)", "").replace("(bool)", "")
        out.append(v)
    w.writerow(out)
