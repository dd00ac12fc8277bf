"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i`) into a small text file for profiles/."""
import csv, subprocess, sys, io

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of {rep.split('/')[-1]}", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"## kernel: {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
        for h, u, v in zip(hdr, units, r):
            name = h.split(" ")[0]
            if any(name == k or name.endswith("." + k) or k in name for k in KEYS) and v not in ("", "n/a"):
                lines.append(f"{name} = {v} {u}")
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        def f(r, k):
            try: return float(r[idx[k]])
            except Exception: return 0.0
        data = rows[2:]
        tot = sum(f(r, "# Samples") for r in data) or 1.0
        agg = sorted(((s, sum(f(r, s) for r in data)) for s in stalls), key=lambda kv: -kv[1])[:8]
        lines.append("## warp stall sampling (all warps, share of samples)")
        for s, v in agg:
            lines.append(f"{s} = {100 * v / tot:.1f} %")
        lines.append("")
        lines.append("## top SASS instructions by samples")
        for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:15]:
            lines.append(f"{int(f(r, '# Samples')):8d}  {r[idx['Source']][:100]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
