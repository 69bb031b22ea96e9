"""Developer tool: trimmed extract of an Nsight Compute report (the raw page of `ncu -i X.ncu-rep --page raw --csv`)
with the metrics profiles/*_summary.md quotes, one row per captured launch.
    ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv > /tmp/raw.csv && python tools/ncu_extract.py /tmp/raw.csv > profiles/rNN_ncu_step_raw.csv"""
import csv
import sys

KEEP = [
    "ID", "Kernel Name", "Grid Size", "Block Size",
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_utcmma.sum",
    "smsp__sass_inst_executed_op_tma_ld.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(k, hdr.index(k)) for k in KEEP if k in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([k for k, _ in cols])
    w.writerow([units[i] for _, i in cols])
    for r in data:
        w.writerow([r[i] for _, i in cols])


if __name__ == "__main__":
    main()
