"""Hot SASS lines of an `ncu --page source --csv` export.   python tools/ncu_hot.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
si = hdr.index("# Samples"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in data)
print("total samples", tot)
idx = sorted(range(len(data)), key=lambda k: -int(data[k][si] or 0))[:top]
for k in sorted(idx):
    r = data[k]
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    print("%5d %6d %5.1f%% ex=%-8s %-70s %s" % (k, int(r[si]), 100.0 * int(r[si]) / tot, r[ex], r[src][:70], " ".join("%s:%d" % (n, c) for c, n in st if c)))
