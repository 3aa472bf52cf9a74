"""Attribute the executed instructions / shared-memory wavefronts / stall samples of one kernel in an
.ncu-rep (captured with --import-source on) to the code between its barriers.
usage: python tools/ncu_segments.py rep.ncu-rep kernel_regex"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
H = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    try:
        return float(r[H[k]] or 0)
    except ValueError:
        return 0.0


seg, acc = 0, {}
for r in rows[hi + 1:]:
    if len(r) < 8 or r[0] == "Address" or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name":
            break
        continue
    ins = r[H["Source"]].strip()
    a = acc.setdefault(seg, [0, 0, 0, 0, {}])
    ie = num(r, "Instructions Executed")
    a[0] += ie
    a[1] += num(r, "L1 Wavefronts Shared")
    a[2] += num(r, "L1 Wavefronts Shared Ideal")
    a[3] += num(r, "# Samples")
    parts = ins.split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    a[4][op] = a[4].get(op, 0) + ie
    if "BAR.SYNC" in ins:
        seg += 1
tot = sum(a[0] for a in acc.values()) or 1
for s, a in acc.items():
    top = sorted(a[4].items(), key=lambda x: -x[1])[:9]
    print(f"seg {s}: inst {a[0] / 1e6:8.1f} M ({100 * a[0] / tot:4.1f}%)  shared wavefronts {a[1] / 1e6:7.1f} M (ideal {a[2] / 1e6:7.1f} M)"
          f"  samples {a[3]:7.0f}   " + " ".join(f"{k}:{v / 1e6:.0f}" for k, v in top))
