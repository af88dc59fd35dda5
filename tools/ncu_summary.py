"""Print the handful of ncu raw-page metrics we track (reads an .ncu-rep via `ncu -i`)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:60])
    for k in keys:
        if k in d: print("  %-75s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = sorted(((float(d[h].replace(",", "")), h.split("stalled_")[1].split("_per_issue")[0]) for h in stall if d[h]), reverse=True)
    print("  stalls/issue:", ", ".join("%s=%.2f" % (n, v) for v, n in st[:8]))
