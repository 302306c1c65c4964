"""Turn the ncu CSV exports a gpurun call brings back into the small, tracked summaries under profiles/.

    python scripts/ncu_summarize.py launches <launches.csv> <out.txt> [<out_traffic.json>]
        per-kernel totals of a `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list
    python scripts/ncu_summarize.py raw <prof_raw.csv> <out.txt>
        the roofline-relevant counters of every launch in a `--set full` capture (--page raw --csv)
"""
import collections
import csv
import json
import re
import sys


def short_name(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*", "", n)
    n = n.replace("efgb::", "").replace("at::native::", "native::")
    return n


def launches(path, out_txt, out_json=None):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    per = collections.OrderedDict()  # launch id -> dict
    for r in data:
        if len(r) <= vi:
            continue
        d = per.setdefault(r[0], {"name": short_name(r[ki])})
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        if r[mi].startswith("gpu__time_duration"):
            d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(u, 1e-6)
        elif r[mi].startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            d["dram"] = d.get("dram", 0.0) + v * mult
    agg = collections.defaultdict(lambda: {"n": 0, "ms": 0.0, "dram": 0.0})
    for d in per.values():
        key = re.sub(r"<.*", "", d["name"])
        a = agg[key]
        a["n"] += 1
        a["ms"] += d.get("ms", 0.0)
        a["dram"] += d.get("dram", 0.0)
    tot = sum(a["ms"] for a in agg.values())
    have_dram = any(a["dram"] for a in agg.values())
    with open(out_txt, "w") as f:
        f.write("ncu launch list of ONE Voxel-DETR training step (bench.py --profile-step, cudaProfilerStart/Stop around the step)\n")
        f.write("serialised, cold-cache per-launch device times: compare SHARES, not absolutes\n")
        f.write("total kernel time %.2f ms over %d launches\n\n" % (tot, len(per)))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:60]:
            line = "%-72s n=%5d %9.3f ms %5.1f%%" % (k[:72], a["n"], a["ms"], 100 * a["ms"] / tot)
            if have_dram:
                line += "  dram %9.1f MB (%.1f MB/launch)" % (a["dram"] / 1e6, a["dram"] / 1e6 / a["n"])
            f.write(line + "\n")
    if out_json:
        js = {k: {"launches": a["n"], "ms": round(a["ms"], 4), "share": round(a["ms"] / tot, 4),
                  "dram_bytes": int(a["dram"]), "dram_bytes_per_launch": int(a["dram"] / a["n"])}
              for k, a in agg.items() if a["ms"] / tot > 0.002}
        json.dump({"source": path, "total_ms": round(tot, 3), "launches": len(per), "kernels": js}, open(out_json, "w"), indent=1)


WANT = [
    ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def raw(path, out_txt):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out_txt, "w") as f:
        f.write("ncu --set full --clock-control none (page raw), one block per captured launch; source: %s\n" % path)
        for r in data:
            f.write("\n%s\n" % short_name(r[idx["Kernel Name"]])[:110])
            for m, label in WANT:
                if m in idx:
                    f.write("  %-12s %s %s\n" % (label, r[idx[m]], units[idx[m]]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        raw(sys.argv[2], sys.argv[3])
