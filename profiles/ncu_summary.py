import csv, subprocess, sys
def summ(rep, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
            "launch__block_size", "launch__cluster_size", "launch__grid_size", "launch__registers_per_thread", "sm__cycles_elapsed.max"]
    out = ["== %s" % title]
    d = dict((h, (u, v)) for h, u, v in zip(hdr, units, vals))
    for k in keys:
        m = [h for h in hdr if h == k]
        if m: out.append("%-86s %s %s" % (k, d[m[0]][1], d[m[0]][0]))
    st = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
            try: st.append((float(d[h][1]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
            except ValueError: pass
    out.append("warp stall reasons (warps stalled per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
    return "\n".join(out)
print(summ(sys.argv[1], sys.argv[2]))
