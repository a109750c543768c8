#!/usr/bin/env python
"""Turns the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py gpurun_out/launches_r01.csv gpurun_out/prof_score_XXX.ncu-rep profiles/ncu_score_r01.md
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.avg"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr:
            d = dict(zip(hdr, r))
            agg.setdefault(d["Kernel Name"], []).append(float(d["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | avg us | share of GPU time |", "|---|---|---|---|"]
    for k, v in agg.items():
        out.append("| `%s` | %d | %.1f | %.1f %% |" % (k[:90], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    return out


def raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = ["kernel: `%s`" % vals[hdr.index("Kernel Name")], "", "| metric | value | unit |", "|---|---|---|"]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append("| %s | %s | %s |" % (w, vals[i], units[i]))
    st = {}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(vals[i])
            except ValueError:
                pass
    try:        # the bench's roofline.traffic reads this
        import json
        import os
        rd, wr = float(vals[hdr.index("dram__bytes_read.sum")]), float(vals[hdr.index("dram__bytes_write.sum")])
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd *= mult[units[hdr.index("dram__bytes_read.sum")]]
        wr *= mult[units[hdr.index("dram__bytes_write.sum")]]
        json.dump({"kernel": vals[hdr.index("Kernel Name")], "dram_bytes_read": rd, "dram_bytes_write": wr, "source": path},
                  open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_score_traffic.json"), "w"))
    except Exception as e:
        print("traffic json not written:", e)
    tot = sum(st.values()) or 1
    out += ["", "warp stall sampling: " + ", ".join("%s %.1f %%" % (k, 100 * v / tot)
                                                    for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8])]
    return out


if __name__ == "__main__":
    launch_csv, rep, dest = sys.argv[1:4]
    lines = ["# ncu summary (round 1)", "", "## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; "
             "per-launch times are cold-cache and serialised: compare shares)", ""] + launches(launch_csv)
    lines += ["", "## `ncu --set full --clock-control none` of the scoring kernel", ""] + raw(rep)
    open(dest, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
