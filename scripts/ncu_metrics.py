"""Print the handful of ncu metrics that matter (roofline inputs + stall mix) for every kernel in a .ncu-rep."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum",
        "smsp__inst_executed_pipe_lsu.sum", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("====", r[idx["Kernel Name"]][:60], "grid", r[idx["launch__grid_size"]], "block", r[idx["launch__block_size"]])
    for w in want:
        if w in idx:
            print(f"  {w:72s} {r[idx[w]]:>16s} {units[idx[w]]}")
    st = sorted(((float(r[idx[s]] or 0), s.split("issue_stalled_")[1].split("_per_issue")[0]) for s in stalls), reverse=True)
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
