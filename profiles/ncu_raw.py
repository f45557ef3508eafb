"""Per-kernel table from an `ncu --set full` report (read on the CPU box):
    python profiles/ncu_raw.py gpurun_out/prof.ncu-rep > profiles/ncu_rNN.txt
duration, DRAM bytes read+written (the `traffic` of bench.py's roofline), DRAM throughput %, L2 bytes, achieved occupancy,
registers, grid size.  Durations are cold-cache, serialised, under the profiler: evidence of traffic/shape, not bench values."""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_bytes.sum", "L2MB"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "blk")]


def main(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    idx = [(hdr.index(c), n) for c, n in COLS if c in hdr]
    print("%-44s " % "kernel" + " ".join("%9s" % n for _, n in idx))
    for r in rows[2:]:
        name = r[ki].replace("void ", "").replace("<unnamed>::", "").split("(")[0][:44]
        vals = []
        for i, n in idx:
            v = r[i].replace(",", "")
            try:
                f = float(v)
                u = units[i]
                if n == "us":
                    f = f / 1000 if u == "ns" else (f * 1000 if u == "ms" else f)
                if n.endswith("MB"):
                    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0) * f
                vals.append("%9.2f" % f if f < 1e5 else "%9.0f" % f)
            except ValueError:
                vals.append("%9s" % v[:9])
        print("%-44s " % name + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
