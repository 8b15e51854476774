"""Summarise an ncu report's SASS page: python tools/ncu_regions.py report.ncu-rep [rows-per-block]
Prints totals and, per block of SASS rows, the share of issued instructions and of stall samples."""
import csv, subprocess, sys
rep = sys.argv[1]; blk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], rows[hi + 1:]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[isamp]) for r in data)
print("kernel:", rows[0][1][:100]); print("warp instructions", tot, "samples", tots, "sass rows", len(data))
for s in range(0, len(data), blk):
    b = data[s:s + blk]
    sm = sum(int(r[isamp]) for r in b); ins = sum(int(r[ia]) for r in b)
    if sm > 0.005 * tots:
        top = max(b, key=lambda r: int(r[isamp]))
        print("rows %4d-%4d samp%%=%5.1f inst%%=%5.1f exec=%9d | top: %s (%d)" % (s, s + blk - 1, 100 * sm / tots, 100 * ins / tot, int(b[0][ia]), top[isrc].strip()[:56], int(top[isamp])))
