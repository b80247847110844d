"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_dsp_tuned.txt [--json profiles/dsp_traffic.json]
"""
import csv, io, json, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
]

def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines, last = [f"# ncu --set full summary of {rep}", ""], None
    for n, r in enumerate(rows[2:]):
        lines.append(f"launch {n}: {r[ix['Kernel Name']]}")
        for k in KEYS:
            if k in ix:
                lines.append(f"  {k:75s} {r[ix[k]]} {units[ix[k]]}")
        lines.append("")
        last = r
    open(out, "w").write("\n".join(lines))
    if "--json" in sys.argv and last is not None:
        def num(k):
            v, u = float(last[ix[k]].replace(",", "")), units[ix[k]]
            return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        tot = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        json.dump({"dram_bytes_per_launch": tot, "kernel": last[ix["Kernel Name"]], "source": rep},
                  open(sys.argv[sys.argv.index("--json") + 1], "w"))
    print("\n".join(lines[:40]))

main()
