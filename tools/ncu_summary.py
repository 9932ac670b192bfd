"""Turn an .ncu-rep into a compact text summary (key raw metrics + top stall sites) for profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## kernel: {name[:110]}\n")
            for h, u, v in zip(hdr, units, r):
                if h in KEYS:
                    f.write(f"{h:70s} {v:>18s} {u}\n")
            f.write("\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        if len(srows) > 2:
            sh = srows[1]
            ix = {h: i for i, h in enumerate(sh)}
            data = [r for r in srows[2:] if len(r) >= len(sh) and r[ix["# Samples"]].isdigit()]
            tot = sum(int(r[ix["# Samples"]]) for r in data)
            f.write(f"## top stall sites of the first kernel (warp-state samples, total {tot})\n")
            top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:25]
            for r in top:
                f.write(f"{int(r[ix['# Samples']]):8d}  exec {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']].strip()[:90]}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
