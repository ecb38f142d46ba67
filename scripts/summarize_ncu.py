#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) of the solve kernel into profiles/<name>.md + .json (run in the build container).

    python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01b_scan_lines "note"
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    summ = []
    for data in rows[2:]:
        d = dict(zip(hdr, data))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name", "?"), "metrics": {}, "stalls": {}}
        for k in KEYS:
            if k in d:
                rec["metrics"][k] = [d[k], u[k]]
        for k in hdr:
            if k.startswith(STALLS) and not k.endswith("_not_issued"):
                try:
                    rec["stalls"][k[len(STALLS):]] = float(d[k].replace(",", ""))
                except ValueError:
                    pass
        summ.append(rec)
    json.dump({"report": rep, "note": note, "launches": summ}, open(out + ".json", "w"), indent=1)
    if len(sys.argv) > 5:
        # top_kernel.json for bench.py: python summarize_ncu.py rep out note <points in the profiled launch> <top_kernel.json>
        pts = float(sys.argv[4])
        m = summ[0]["metrics"]

        def val(k):
            v, u = m[k]
            v = float(v.replace(",", ""))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
        top = {"profile": out + ".md", "kernel": summ[0]["kernel"][:60],
               "dram_bytes_per_point": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / pts,
               "fp64_pipe_util_pct": float(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]),
               "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
               "points_in_profiled_launch": pts}
        json.dump(top, open(sys.argv[5], "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("# ncu summary: %s\n\n%s\n\n" % (rep, note))
        for rec in summ:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % rec["kernel"])
            for k, (v, un) in rec["metrics"].items():
                f.write("| %s | %s | %s |\n" % (k, v, un))
            tot = sum(rec["stalls"].values()) or 1.0
            f.write("\nWarp-state samples (pc sampling):\n\n| state | samples | share |\n|---|---|---|\n")
            for k, v in sorted(rec["stalls"].items(), key=lambda kv: -kv[1])[:10]:
                f.write("| %s | %d | %.1f%% |\n" % (k, v, 100 * v / tot))
            f.write("\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
