"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, sys, collections, re

def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, v * scale))
    tot = sum(t for _, t in rows)
    agg = collections.OrderedDict()
    for n, t in rows:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
    lines = [f"# {path}: {len(rows)} launches, {tot:.1f} us total (cold-cache, serialised under ncu: compare shares)",
             f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{n[:60]:60s} {c:8d} {t:12.1f} {t / c:10.2f} {t / tot:7.3f}")
    txt = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(txt)
    print(txt)

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
