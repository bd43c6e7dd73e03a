#!/usr/bin/env python
"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into tracked files under profiles/:

    python tools/ncu_summarise.py REPORT.ncu-rep LABEL [--traffic kernel_key:rows:f:k:mode ...]

writes profiles/LABEL_summary.csv (one line per captured launch: duration, DRAM bytes, L2 sectors, tensor /
XU pipe activity, registers, top stall reasons) and, with --traffic, adds / replaces entries in
profiles/ncu_traffic.json -- the per-launch dram__bytes_read.sum + dram__bytes_write.sum that bench.py
reports as roofline.traffic.  kernel_key is bench.py's phase name (ratio, coefficient, numerator, fused,
sparse_rows, sparse_scatter) and is matched against the demangled kernel name by the table below.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum"]
MATCH = {"ratio": "tc_gemm_kernel<256, 0, 1, 0, 1", "coefficient": "tc_gemm_kernel<256, 0, 0, 0, 0",
         "numerator": "tc_gemm_kernel<256, 1, 1, 0, 0", "fused": "fused_coef_kernel", "fused256": "fused_coef256_kernel",
         "fp64_ratio": "generic_gemm_kernel<double, 1>", "fp64_coefficient": "generic_gemm_kernel<double, 2>",
         "fp64_numerator": "generic_gemm_kernel<double, 3>",
         "sparse_rows": "sparse_rows_kernel", "sparse_scatter": "sparse_numerator_bcsc_kernel"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep, label = sys.argv[1], sys.argv[2]
    traffic = []
    if "--traffic" in sys.argv:
        traffic = [t.split(":") for t in sys.argv[sys.argv.index("--traffic") + 1:]]
    hdr, units, body = raw_page(rep)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    path = os.path.join(ROOT, "profiles", label + "_summary.csv")
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h]
    with open(path, "w") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel"] + ["%s [%s]" % (k, units[hdr.index(k)]) for k in KEYS if k in hdr] + ["top_stalls"])
        for r in body:
            name = r[hdr.index("Kernel Name")]
            st = sorted(((float(r[hdr.index(h)] or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stall),
                        reverse=True)[:5]
            w.writerow([name[:140]] + [r[hdr.index(k)] for k in KEYS if k in hdr] +
                       [" ".join("%s=%d" % (n, v) for v, n in st)])
    print("wrote", path)
    if traffic:
        tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        entries = json.load(open(tj)) if os.path.exists(tj) else []
        for key, rows_, f, k, mode in traffic:
            if key == "sparse":      # one sparse iteration = rows pass + scatter pass: sum the first launch of each
                rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                tr = tw = 0.0
                for sub in ("sparse_rows_", "sparse_numerator_bcsc_kernel"):   # sparse_rows_kernel or sparse_rows_full_kernel
                    for r in body:
                        if sub in r[hdr.index("Kernel Name")]:
                            tr += float(r[rd]) * UNIT[units[rd]]
                            tw += float(r[wr]) * UNIT[units[wr]]
                            break
                e = {"kernel": key, "rows": int(rows_), "f": int(f), "k": int(k), "mode": mode, "dram_bytes": tr + tw,
                     "dram_bytes_read": tr, "dram_bytes_write": tw, "report": os.path.basename(rep),
                     "kernel_name": "sparse_rows_(full_)kernel + sparse_numerator_bcsc_kernel"}
                entries = [x for x in entries if not (x["kernel"] == key and x["rows"] == e["rows"] and x["f"] == e["f"]
                                                      and x["k"] == e["k"] and x["mode"] == mode)]
                entries.append(e)
                continue
            for r in body:
                if MATCH[key] in r[hdr.index("Kernel Name")]:
                    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                    tot = float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]]
                    key = key[5:] if key.startswith("fp64_") else key      # bench.py looks the phase name up, the mode tells them apart
                    e = {"kernel": key, "rows": int(rows_), "f": int(f), "k": int(k), "mode": mode, "dram_bytes": tot,
                         "dram_bytes_read": float(r[rd]) * UNIT[units[rd]], "dram_bytes_write": float(r[wr]) * UNIT[units[wr]],
                         "report": os.path.basename(rep), "kernel_name": r[hdr.index("Kernel Name")][:120]}
                    entries = [x for x in entries if not (x["kernel"] == key and x["rows"] == e["rows"] and x["f"] == e["f"]
                                                          and x["k"] == e["k"] and x["mode"] == mode)]
                    entries.append(e)
                    break
        json.dump(entries, open(tj, "w"), indent=1)
        print("updated", tj)


if __name__ == "__main__":
    main()
