"""ncu report -> profiles/<name>.csv (selected metrics, one column per kernel) + profiles/kernel_metrics.json.

  python tools/ncu_extract.py gpurun_out/prof.ncu-rep profiles/r2_ncu_full_step.csv

bench.py reads kernel_metrics.json by kernel name for `roofline.traffic` and the kNN block; every number in it comes from
the report named in its "source" field (per launch, first launch of each kernel)."""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kcol = col["Kernel Name"]
    seen, kernels = set(), []
    for r in data:
        if r[kcol] not in seen:
            seen.add(r[kcol]); kernels.append(r)
    metrics = [h for h in hdr if h in KEEP or h.startswith(STALLS)]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[kcol][:70] for r in kernels])
        for m in metrics:
            w.writerow([m, units[col[m]]] + [r[col[m]] for r in kernels])

    def num(r, m, scale=1.0):
        try:
            return float(r[col[m]].replace(",", "")) * scale
        except Exception:
            return None

    def unit_scale(m):
        u = units[col[m]].lower()
        return {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
    table = {}
    for r in kernels:
        rd = num(r, "dram__bytes_read.sum", unit_scale("dram__bytes_read.sum")) or 0.0
        wr = num(r, "dram__bytes_write.sum", unit_scale("dram__bytes_write.sum")) or 0.0
        table[r[kcol]] = {
            "dram_bytes": rd + wr,
            "duration_us": num(r, "gpu__time_duration.sum", unit_scale("gpu__time_duration.sum")),
            "inst_executed": num(r, "smsp__inst_executed.sum"),
            "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "xu_pipe_pct": num(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "tensor_active_pct": num(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
            "registers": num(r, "launch__registers_per_thread"),
        }
    dst = os.path.join(os.path.dirname(os.path.abspath(out_csv)), "kernel_metrics.json")
    with open(dst, "w") as f:
        json.dump({"source": os.path.relpath(out_csv, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                   "kernels": table}, f, indent=1)
    print(f"{len(kernels)} kernels, {len(metrics)} metrics -> {out_csv}, {dst}")


if __name__ == "__main__":
    main()
