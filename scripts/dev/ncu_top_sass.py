"""Top sampled SASS instructions per kernel of an `ncu --page source --csv` dump.  usage: ncu_top_sass.py dump.csv [kernel substring] [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for s in secs:
    if want not in s["name"] or s["name"] in seen:
        continue
    seen.add(s["name"])
    hdr = s["rows"][0]; data = [r for r in s["rows"][1:] if len(r) == len(hdr)]
    iS, iSrc, iE = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[iS] or 0) for r in data)
    print("==", s["name"][:90], "samples", tot, "instrs", len(data))
    top = sorted(enumerate(data), key=lambda t: -int(t[1][iS] or 0))[:n]
    for idx, r in sorted(top):
        print(f"{idx:5d} {100.0 * int(r[iS] or 0) / max(tot, 1):5.1f}% {r[iE]:>9s}  {r[iSrc][:120]}")
