"""Summarise an .ncu-rep (read here with `ncu -i … --page raw --csv`) into the small CSVs committed under profiles/.
  python scripts/ncu_summary.py step  <rep> <out.csv> [kernel-substring]   -> metric,unit,launch0 rows for the first matching launch
  python scripts/ncu_summary.py table <rep> <out.csv>                      -> one row per launch with the tensor-core / DRAM columns
"""
import csv
import io
import subprocess
import sys

STEP_METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
                "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
                "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "launch__occupancy_limit_registers"]
TABLE_METRICS = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                 "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def main():
    mode, rep, outp = sys.argv[1], sys.argv[2], sys.argv[3]
    hdr, units, data = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    if mode == "step":
        sub = sys.argv[4] if len(sys.argv) > 4 else "b2q_step_kernel"
        row = next(r for r in data if sub in r[ix["Kernel Name"]])
        with open(outp, "w") as f:
            f.write("metric,unit,launch0\n")
            f.write('Kernel Name,,"%s"\n' % row[ix["Kernel Name"]])
            names = STEP_METRICS + sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))
            for m in names:
                if m in ix:
                    f.write("%s,%s,%s\n" % (m, units[ix[m]], row[ix[m]].replace(",", "")))
    else:
        with open(outp, "w") as f:
            cols = [m for m in TABLE_METRICS if m in ix]
            f.write("Kernel Name," + ",".join("%s [%s]" % (m, units[ix[m]]) for m in cols) + "\n")
            for r in data:
                name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
                f.write('"%s",' % name + ",".join(r[ix[m]].replace(",", "") for m in cols) + "\n")


if __name__ == "__main__":
    main()
