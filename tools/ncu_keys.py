"""Print a compact table of key metrics for every launch in an .ncu-rep (ncu --set full capture).

    python tools/ncu_keys.py gpurun_out/prof.ncu-rep [extra_metric_substring ...]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "sm__cycles_active.avg", "smsp__inst_executed.sum",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("====", r[hdr.index("ID")], r[name_i][:70])
        for i, h in enumerate(hdr):
            hh = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[0].isupper() else h
            if h in KEYS or hh in KEYS or any(e in h for e in extra):
                print(f"  {h:90s} {units[i]:12s} {r[i]}")


if __name__ == "__main__":
    main()
