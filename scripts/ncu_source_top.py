"""Summarise an `ncu --page source --csv` dump: per kernel, the instructions with the most stall samples."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern = None; hdr = None; rows = []; out = []
def flush():
    if kern and rows:
        out.append((kern, hdr, list(rows)))
for r in csv.reader(open(path)):
    if r and r[0] == "Kernel Name":
        flush(); kern = r[1]; hdr = None; rows = []
    elif r and r[0] == "Address":
        hdr = r
    elif hdr and len(r) >= len(hdr) - 1:
        rows.append(r)
flush()
for ki, (kern, hdr, rows) in enumerate(out):
    si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in rows)
    insts = sum(int(r[ie]) for r in rows)
    print("=" * 100); print("#%d %s  samples=%d  warp-insts=%d" % (ki, kern[:80], tot, insts))
    agg = {hdr[i]: sum(int(r[i] or 0) for r in rows) for i in stall_cols}
    print("  stalls:", ", ".join("%s=%.0f%%" % (k[6:], 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    for n, r in sorted(((int(r[si]), r) for r in rows), key=lambda t: -t[0])[:topn]:
        top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        idx = rows.index(r)
        print("  %5.1f%%  [%4d] %-70s x%-8s %s" % (100 * n / max(tot, 1), idx, r[src].strip()[:70], r[ie], " ".join("%s:%d" % (b, a) for a, b in top)))
