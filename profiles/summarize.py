"""Turn an .ncu-rep (brought back in gpurun_out/) into the short text summary committed under
profiles/.  Usage: python profiles/summarize.py gpurun_out/X.ncu-rep > profiles/r1_X.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[ki][:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:75s} {r[i]:>18s} {units[i]}")
        st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")
              and r[i].replace(".", "").isdigit()]
        tot = sum(v for v, _ in st) or 1.0
        for v, h in sorted(st, reverse=True)[:6]:
            print(f"  stall {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):40s} {100 * v / tot:5.1f} %")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
