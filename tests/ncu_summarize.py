"""Aggregates an ncu launch list (`--csv --log-file launches.csv`, metrics gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum) per kernel name.   python tests/ncu_summarize.py launches.csv out_prefix [title]
Writes <out_prefix>.json and <out_prefix>.txt (the files bench.py's roofline.traffic and profiles/README.md refer to)."""
import csv
import json
import re
import sys


def to_bytes(v, unit):
    unit = unit.strip().lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "b": 1}.get(unit, 1)
    return float(v.replace(",", "")) * mult


def to_us(v, unit):
    unit = unit.strip().lower()
    mult = {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(unit, 1)
    return float(v.replace(",", "")) * mult


def main():
    path, prefix = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else "ncu launch list"
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = {}
    launches = {}
    for r in rows:
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("lr::", "").strip()
        key = (r["ID"], name)
        d = launches.setdefault(key, {"us": 0.0, "rd": 0.0, "wr": 0.0})
        m, v, u = r["Metric Name"], r["Metric Value"], r["Metric Unit"]
        if m == "gpu__time_duration.sum":
            d["us"] = to_us(v, u)
        elif m == "dram__bytes_read.sum":
            d["rd"] = to_bytes(v, u)
        elif m == "dram__bytes_write.sum":
            d["wr"] = to_bytes(v, u)
    for (_, name), d in launches.items():
        a = per.setdefault(name, {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        a["launches"] += 1
        a["us"] += d["us"]
        a["dram_read"] += d["rd"]
        a["dram_write"] += d["wr"]
    json.dump(per, open(prefix + ".json", "w"), indent=1)
    tot = sum(a["us"] for a in per.values())
    out = [f"{title} - cold-cache serialised per-launch times; compare SHARES",
           f"{'kernel':44s} {'launches':>8s} {'total us':>10s} {'share':>6s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'MB/launch':>10s}"]
    for name, a in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        out.append(f"{name[:44]:44s} {a['launches']:8d} {a['us']:10.1f} {100 * a['us'] / tot:5.1f}% "
                   f"{a['dram_read'] / 1e6:11.1f} {a['dram_write'] / 1e6:11.1f} "
                   f"{(a['dram_read'] + a['dram_write']) / 1e6 / a['launches']:10.2f}")
    out.append(f"{'TOTAL':44s} {sum(a['launches'] for a in per.values()):8d} {tot:10.1f}")
    open(prefix + ".txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
