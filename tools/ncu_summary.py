"""tools/ncu_summary.py — turns the ncu CSV pages tools/ncu_capture.sh brings back into the text summaries under profiles/."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ("Duration", "DRAM Throughput", "Memory Throughput", "L2 Cache Throughput", "Compute (SM) Throughput", "Registers Per Thread", "Achieved Occupancy",
        "Theoretical Occupancy", "Issue Slots Busy", "Executed Ipc Active", "L2 Hit Rate", "L1/TEX Hit Rate", "Mem Busy", "Max Bandwidth",
        "Warp Cycles Per Issued Instruction", "No Eligible", "Eligible Warps Per Scheduler", "Avg. Active Threads Per Warp",
        "Avg. Not Predicated Off Threads Per Warp", "Branch Efficiency", "Grid Size", "Block Size", "Static Shared Memory Per Block", "Waves Per SM")
RAW = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
       "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")


def details(fn):
    rows = list(csv.reader(open(fn)))
    hi = [i for i, r in enumerate(rows) if "Metric Name" in r][0]
    h = rows[hi]
    mi, vi, ui, si, ki = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Section Name"), h.index("Kernel Name")
    return rows[hi + 1][ki], [(r[si], r[mi], r[ui], r[vi]) for r in rows[hi + 1:] if len(r) > vi]


def raw(fn):
    rows = list(csv.reader(open(fn)))
    hi = [i for i, r in enumerate(rows) if "ID" in r and "Kernel Name" in r][0]
    h, u, v = rows[hi], rows[hi + 1], rows[hi + 2]
    return {h[i]: (v[i], u[i]) for i in range(len(h))}


def main():
    d = os.path.join(ROOT, "profiles", "r02_ncu")
    notes = json.load(open(os.path.join(d, "notes.json"))) if os.path.exists(os.path.join(d, "notes.json")) else {}
    for fn in sorted(os.listdir(d)):
        if not fn.endswith("_details.csv"):
            continue
        name = fn[:-len("_details.csv")]
        kern, det = details(os.path.join(d, fn))
        rw = raw(os.path.join(d, name + "_raw.csv"))
        out = ["ncu --set full --clock-control none --import-source on  (tools/ncu_capture.sh; CSV pages: profiles/r02_ncu/%s_{details,raw}.csv)" % name,
               "kernel: " + kern, notes.get(name, ""), ""]
        for s, m, u, v in det:
            if m in WANT:
                out.append("  %-34s %-44s %12s %s" % (s[:34], m, v, u))
        out.append("")
        out.append("  stall reasons (warps per issue-active cycle, top 8):")
        st = sorted(((float(v[0].replace(",", "")) if v[0] not in ("", "n/a") else 0.0, k) for k, v in rw.items()
                     if "issue_stalled" in k and "average" in k and k.endswith(".ratio")), reverse=True)[:8]
        for val, k in st:
            out.append("    %8.2f  %s" % (val, k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        out.append("")
        for k in RAW:
            if k in rw:
                out.append("  %-52s %s %s" % (k, rw[k][0], rw[k][1]))
        open(os.path.join(ROOT, "profiles", name + "_summary.txt"), "w").write("\n".join(out) + "\n")
        print("wrote", name)


if __name__ == "__main__":
    main()
