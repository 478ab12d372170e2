"""Reduce `ncu -i X.ncu-rep --page raw --csv` (2000+ columns) to the columns the profiles/ tables quote.

usage: ncu_raw_table.py RAW.csv [--csv OUT.csv] [--title "..."] > table.md

Durations are normalised to microseconds and byte counts to MB whatever unit ncu chose per column.
`tensor %` = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active (the tensor-pipe-active figure north_star asks for),
`dram MB` = dram__bytes_read.sum + dram__bytes_write.sum for the launch (what bench.py's roofline.traffic quotes).
"""
import argparse, csv, re, sys

COLS = [
    ("Kernel Name", "kernel", None),
    ("Grid Size", "grid", None),
    ("Block Size", "block", None),
    ("launch__registers_per_thread", "regs", None),
    ("gpu__time_duration.sum", "us", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", None),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", None),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", None),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %", None),
    ("dram__bytes_read.sum", "dram rd MB", "bytes"),
    ("dram__bytes_write.sum", "dram wr MB", "bytes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", None),
    ("smsp__cycles_active.avg", "active cyc", None),
]
TIME = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}
BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "B": 1e-6, "KB": 1e-3, "MB": 1.0, "GB": 1e3}


def short(name):
    name = name.replace("<unnamed>::", "").replace("void ", "")
    m = re.match(r"([\w:]+(<[^(]*>)?)", name)
    return (m.group(1) if m else name)[:80]


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {}
        for col, label, kind in COLS:
            if col not in hdr:
                rec[label] = ""
                continue
            i = hdr.index(col)
            v = r[i]
            if kind == "time":
                v = float(v.replace(",", "")) * TIME.get(units[i], 1.0)
            elif kind == "bytes":
                v = float(v.replace(",", "")) * BYTES.get(units[i], 1e-6)
            elif label == "kernel":
                v = short(v)
            elif label not in ("grid", "block"):
                try:
                    v = float(v.replace(",", ""))
                except ValueError:
                    pass
            rec[label] = v
        out.append(rec)
    return out


def fmt(v):
    if isinstance(v, float):
        return f"{v:.2f}" if abs(v) < 1000 else f"{v:.0f}"
    return str(v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("--csv")
    ap.add_argument("--title", default="")
    a = ap.parse_args()
    recs = load(a.raw)
    labels = [l for _, l, _ in COLS]
    if a.csv:
        w = csv.writer(open(a.csv, "w", newline=""))
        w.writerow(labels)
        for r in recs:
            w.writerow([fmt(r[l]) for l in labels])
    if a.title:
        print(f"### {a.title}\n")
    print("| # | " + " | ".join(labels) + " |")
    print("|---|" + "---|" * len(labels))
    for i, r in enumerate(recs):
        print(f"| {i} | " + " | ".join((f"`{r[l]}`" if l == "kernel" else fmt(r[l])) for l in labels) + " |")
    tot = sum(r["us"] for r in recs if isinstance(r["us"], float))
    print(f"\n{len(recs)} launches, {tot:.1f} us summed (each launch replayed alone under ncu: cold caches, no overlap).")


if __name__ == "__main__":
    main()
