#!/usr/bin/env python3
"""Summarise an `ncu --set full` capture of the non-GEMM kernels of one network pass into profiles/*.json.

  ncu --set full --clock-control none --import-source on \
      -k regex:"first_conv_gn|upsample2x_tiled|attention_mma|gn_image16|step_kernel" -s 7 -c 7 \
      -o gpurun_out/prof_misc python tools/ncu_target.py 128 1
  python tools/ncu_misc_summary.py gpurun_out/prof_misc.ncu-rep profiles/r1_misc_ncu.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {"us": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pipe_active_pct": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size"}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    unit = dict(zip(head, units))
    res = []
    for row in data:
        d = dict(zip(head, row))
        e = {"kernel": d["Kernel Name"][:60]}
        for k, m in KEYS.items():
            try:
                v = float(d.get(m, "").replace(",", ""))
            except ValueError:
                continue
            u = unit.get(m, "")
            if k == "us":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
            if k.startswith("dram_r") or k.startswith("dram_w"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                k += "_MB"
            e[k] = round(v, 2)
        if "dram_read_MB" in e and e.get("us"):
            e["dram_GBps"] = round((e["dram_read_MB"] + e["dram_write_MB"]) * 1e-3 / (e["us"] * 1e-6), 0)
        res.append(e)
    json.dump({"source": "ncu --set full --clock-control none, python tools/ncu_target.py 128 1 (256 images per pass); " + rep,
               "kernels": res}, open(out, "w"), indent=1)
    for e in res:
        print(e)


if __name__ == "__main__":
    main()
