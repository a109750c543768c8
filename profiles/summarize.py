"""Turns one `ncu --set full --clock-control none` capture (.ncu-rep) of a kernel into the committed summary files:

    python profiles/summarize.py gpurun_out/prof_score_c2_r02.ncu-rep profiles/ncu_score_r02   [--traffic]

writes <out>.json (the numbers bench.py quotes: pipe-busy percentages, issue-active, DRAM bytes per launch, stall
shares) and <out>.md (the same as a table).  With --traffic it also refreshes profiles/ncu_score_traffic.json, the
`roofline.traffic` source of the bench line.  Needs `ncu` on PATH (reads the report, does not profile)."""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    hdr, units, rows = raw_page(rep)
    row = rows[-1]
    d = dict(zip(hdr, row))
    u = dict(zip(hdr, units))
    res = {"kernel": d.get("Kernel Name"), "source": rep, "metrics": {}}
    for k in KEEP:
        if k in d and d[k] not in ("", None):
            try:
                res["metrics"][k] = {"value": float(d[k].replace(",", "")), "unit": u.get(k, "")}
            except ValueError:
                pass
    stalls = {}
    for k, v in d.items():
        if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v:
            try:
                stalls[k[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(v.replace(",", ""))
            except ValueError:
                pass
    tot = sum(stalls.values())
    if tot:
        res["stall_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
    m = res["metrics"]
    g = lambda k: m[k]["value"] if k in m else None
    res["dmma_pipe_pct"] = g("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active")
    res["fp64_pipe_pct"] = g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
    res["issue_active_pct"] = g("smsp__issue_active.avg.pct_of_peak_sustained_active")

    def nbytes(k):
        if k not in m:
            return None
        return m[k]["value"] * UNIT_SCALE.get(m[k]["unit"], 1.0)
    res["dram_bytes_read"], res["dram_bytes_write"] = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
    with open(out + ".json", "w") as fh:
        json.dump(res, fh, indent=1)
    with open(out + ".md", "w") as fh:
        fh.write("# `ncu --set full --clock-control none`: %s\n\nsource: `%s`\n\n| metric | value | unit |\n|---|---|---|\n"
                 % (res["kernel"], rep))
        for k in KEEP:
            if k in m:
                fh.write("| %s | %s | %s |\n" % (k, m[k]["value"], m[k]["unit"]))
        if "stall_pct" in res:
            fh.write("\nwarp stall sampling: " + ", ".join("%s %.1f %%" % kv for kv in res["stall_pct"].items()) + "\n")
    if "--traffic" in sys.argv:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_score_traffic.json"), "w") as fh:
            json.dump({"kernel": res["kernel"], "dram_bytes_read": res["dram_bytes_read"],
                       "dram_bytes_write": res["dram_bytes_write"], "source": rep}, fh)
    print(json.dumps({k: res[k] for k in ("kernel", "dmma_pipe_pct", "fp64_pipe_pct", "issue_active_pct", "dram_bytes_read")}))


if __name__ == "__main__":
    main()
