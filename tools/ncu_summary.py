"""Summarise an `ncu --set full` capture (raw page) into a small text table for profiles/."""
import csv, subprocess, sys, json
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("dram__bytes_read.sum.per_second", "rd/s"),
        ("dram__bytes_write.sum.per_second", "wr/s"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__cycles_elapsed.max.per_second", "GHz")]
print(" | ".join("%s[%s]" % (short, units[col[name]]) for name, short in want if name in col))
res = []
for r in rows[2:]:
    vals = {short: r[col[name]] for name, short in want if name in col}
    vals["kernel"] = vals["kernel"].split("(")[0][-44:]
    res.append(vals)
    print(" | ".join(str(vals[s])[:44] for _, s in want if s in vals))
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], "w"), indent=1)
