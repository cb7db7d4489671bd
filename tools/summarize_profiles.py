"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/.
  python tools/summarize_profiles.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
  python tools/summarize_profiles.py report gpurun_out/prof_draw_r01.ncu-rep profiles/r01_rfk_draw.md [units_per_launch [kernel_regex]]
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_global_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
]


def ncu_csv(report, *args):
    out = subprocess.run(["ncu", "-i", report, "--csv", *args], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[start]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip()
        d = per.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += float(r[iv].replace(",", "")) / 1e6
    total = sum(v[1] for v in per.values())
    with open(dst, "w") as fh:
        fh.write("# ncu launch list (gpu__time_duration.sum, --clock-control none): per-kernel totals\n\n")
        fh.write("Source: `%s` (per-launch times are cold-cache and serialised: compare shares, not absolutes).\n\n" % src)
        fh.write("| kernel | launches | total ms | mean ms | share |\n|---|---:|---:|---:|---:|\n")
        for name, (n, ms) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.3f | %.4f | %.2f %% |\n" % (name, n, ms, ms / n, 100 * ms / total))
        fh.write("| total | %d | %.3f | | |\n" % (sum(v[0] for v in per.values()), total))


def report(src, dst, units=None, kernel=None):
    pick = ["-k", "regex:" + kernel] if kernel else []
    raw = ncu_csv(src, "--page", "raw", *pick)
    hdr, unit, data = raw[0], raw[1], raw[2:]
    name = data[0][hdr.index("Kernel Name")]
    with open(dst, "w") as fh:
        fh.write("# ncu --set full summary: `%s`\n\nSource report: `%s` (%d launch(es) captured; values of the first).\n\n" % (name, src, len(data)))
        fh.write("| metric | value | unit |\n|---|---:|---|\n")
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = data[0][i]
                fh.write("| %s | %s | %s |\n" % (k, data[0][i], unit[i]))
        if units and "smsp__inst_executed.sum" in vals:
            fh.write("\nWarp instructions per unit of work (%s units per launch): **%.1f**\n" % (units, float(vals["smsp__inst_executed.sum"].replace(",", "")) / float(units)))
        # the counters bench.py quotes (roofline.traffic, roofline.issue), next to the summary
        def num(k):
            return float(vals[k].replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit[hdr.index(k)], 1.0)
        if "dram__bytes_read.sum" in vals and "smsp__inst_executed.sum" in vals:
            import json
            json.dump({"source": src, "kernel": name, "gpu_time_ms": float(vals["gpu__time_duration.sum"]) if "gpu__time_duration.sum" in vals else None,
                       "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                       "warp_instructions_per_launch": num("smsp__inst_executed.sum"), "units_per_launch": float(units) if units else None,
                       "warp_inst_per_unit": num("smsp__inst_executed.sum") / float(units) if units else None,
                       "issue_active_pct": float(vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "nan")),
                       # L2 reduction sectors of the launch and their share of the sustained peak: sectors / time / share = the peak rate
                       "red_sectors_per_launch": num("lts__t_sectors_srcunit_tex_op_red.sum") if "lts__t_sectors_srcunit_tex_op_red.sum" in vals else None,
                       "red_sectors_pct_of_peak": float(vals.get("lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed", "nan").replace(",", ""))},
                      open(dst.rsplit(".", 1)[0] + ".json", "w"), indent=1)
        if "dram__bytes_read.sum" in vals:
            fh.write("\nDRAM traffic per launch: read %s + write %s (%s).\n" % (vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"], unit[hdr.index("dram__bytes_read.sum")]))
        # opcode mix
        rows = ncu_csv(src, "--page", "source", "--print-source", "sass", *pick)
        starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
        if starts:
            h = rows[starts[0]]
            isrc, iex = h.index("Source"), h.index("Instructions Executed")
            end = starts[1] - 1 if len(starts) > 1 else len(rows)
            body = [r for r in rows[starts[0] + 1:end] if len(r) > iex and r[iex].isdigit()]
            total = sum(int(r[iex]) for r in body)
            ops = collections.Counter()
            for r in body:
                m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[isrc])
                ops[m.group(2) if m else "?"] += int(r[iex])
            fh.write("\n## Executed warp instructions by opcode (top 24 of %d)\n\n| opcode | share |%s\n|---|---:|%s\n" % (
                total, " per unit |" if units else "", "---:|" if units else ""))
            for k, v in ops.most_common(24):
                fh.write("| %s | %.2f %% |%s\n" % (k, 100.0 * v / total, (" %.2f |" % (v / float(units))) if units else ""))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None, sys.argv[5] if len(sys.argv) > 5 else None)
