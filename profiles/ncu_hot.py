"""Hot spots of one kernel from an ncu report's source page (needs -lineinfo and --import-source on):
    python profiles/ncu_hot.py prof.ncu-rep <kernel-regex> [launch-skip]
Prints the SASS regions and the instructions with most warp-stall samples."""
import csv
import io
import subprocess
import sys


def main(path, kern, skip="0"):
    txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = [i for i, r in enumerate(rows) if "# Samples" in r]
    hdr = rows[hdr_i[0]]
    end = hdr_i[1] if len(hdr_i) > 1 else len(rows)
    data = [r for r in rows[hdr_i[0] + 1:end] if len(r) == len(hdr)]
    si, ii, src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot = sum(int(r[si]) for r in data)
    print(rows[0][1][:100])
    print("SASS lines %d, samples %d, warp instructions %d" % (len(data), tot, sum(int(r[ii]) for r in data)))
    step = max(len(data) // 24, 1)
    for k in range(0, len(data), step):
        s = sum(int(r[si]) for r in data[k:k + step]); e = sum(int(r[ii]) for r in data[k:k + step])
        print("  [%5d..%5d) samples %6d (%4.1f%%) instr %9d" % (k, k + step, s, 100.0 * s / max(tot, 1), e))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:30]
    for i in sorted(top):
        print("  %5d  samples %5s instr %8s  %s" % (i, data[i][si], data[i][ii], data[i][src][:90]))


if __name__ == "__main__":
    main(*sys.argv[1:])
