"""Turn an .ncu-rep (read here, no GPU needed) into a small markdown summary for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.md [title]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    lines = ["# %s" % title, "",
             "Source: `%s` (ncu --set full --clock-control none; cold-cache, serialised "
             "replays: compare shares, not absolutes).", ""]
    lines[2] = lines[2] % rep
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append("## launch id %s: `%s`" % (d.get("ID"), d.get("Kernel Name", "")[:110]))
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in d:
                lines.append("| %s | %s | %s |" % (k, d[k], u[k]))
        stalls = {k: float(v) for k, v in d.items()
                  if k.startswith("smsp__average_warps_issue_stalled") and
                  k.endswith("per_issue_active.ratio") and v not in ("", "n/a")}
        lines.append("")
        lines.append("Top warp stall reasons (warps stalled per issue-active cycle): " + ", ".join(
            "%s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace(
                "_per_issue_active.ratio", ""), v)
            for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6]))
        lines.append("")
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
