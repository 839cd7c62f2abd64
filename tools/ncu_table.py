"""ncu report (.ncu-rep) -> text table of the metrics DESIGN.md argues from.  usage: ncu_table.py in.ncu-rep out.txt [title]"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
     ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
     ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "registers/thread"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("launch__grid_size", "grid"),
     ("launch__cluster_size", "cluster size")]
with open(out, "w") as f:
    f.write("# %s\n# ncu --set full --clock-control none (one launch each, cold caches, serialised: use shares and ratios, not absolute times)\n\n" % title)
    for r in rows[2:]:
        f.write("%s\n" % r[idx["Kernel Name"]][:110])
        for key, label in M:
            if key in idx and r[idx[key]] != "":
                f.write("    %-26s %s %s\n" % (label, r[idx[key]], units[idx[key]]))
        f.write("\n")
print(open(out).read()[:600])
