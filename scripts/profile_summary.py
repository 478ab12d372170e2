"""Turn ncu outputs into the markdown summaries kept under profiles/.

usage: profile_summary.py --launches LAUNCHES.csv [--rep REPORT.ncu-rep] [--first-kernel gather_batch] --title "..." > profiles/x.md

--launches : CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv` over a bench.py run; the last complete
             step (from one occurrence of --first-kernel to the next) is listed with each kernel's share of the step.
--rep      : `ncu --set full` report; per kernel: duration, DRAM bytes, tensor-pipe / SM / DRAM utilisation, registers, grid.
"""
import argparse, csv, io, re, subprocess, sys


def short(name):
    name = name.replace("<unnamed>::", "").replace("void ", "")
    m = re.match(r"([\w:]+(<[^(]*>)?)", name)
    return (m.group(1) if m else name)[:90]


def launches_table(path, first):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    names = [short(r[4]) for r in rows]
    idx = [i for i, n in enumerate(names) if n.startswith(first)]
    if len(idx) < 2:
        return f"(fewer than two `{first}` launches in {path})\n"
    a, b = idx[-2], idx[-1]
    tot = sum(int(r[-1]) for r in rows[a:b])
    out = [f"{len(rows)} launches profiled; last complete step = launches {a}..{b - 1} ({b - a} kernels, {tot / 1e3:.2f} us summed, "
           "cold-cache and serialised under ncu)\n", "| # | kernel | grid | block | us | share |", "|---|---|---|---|---|---|"]
    for j, (r, n) in enumerate(list(zip(rows, names))[a:b]):
        out.append(f"| {j} | `{n}` | {r[8]} | {r[7]} | {int(r[-1]) / 1e3:.2f} | {100.0 * int(r[-1]) / tot:.1f}% |")
    return "\n".join(out) + "\n"


def rep_table(rep, maxk):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = [("gpu__time_duration.sum", "dur"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
            ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
            ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]
    cols = [(hdr.index(k), lab, units[hdr.index(k)]) for k, lab in want if k in hdr]
    ki = hdr.index("Kernel Name")
    out = ["| kernel | " + " | ".join(f"{lab} ({u})" if u else lab for _, lab, u in cols) + " |", "|---|" + "---|" * len(cols)]
    seen = {}
    gi = hdr.index("launch__grid_size") if "launch__grid_size" in hdr else None
    for r in rows[2:]:
        n = short(r[ki])
        key = (n, r[gi] if gi is not None else "")            # one row per (kernel, grid): the same kernel at another shape is listed too
        seen[key] = seen.get(key, 0) + 1
        if seen[key] > 1 or len(seen) > maxk:
            continue
        vals = []
        for i, _, _ in cols:
            try:
                vals.append(f"{float(r[i].replace(',', '')):.3f}".rstrip("0").rstrip("."))
            except ValueError:
                vals.append(r[i])
        out.append(f"| `{n}` | " + " | ".join(vals) + " |")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--first-kernel", default="gather_batch")
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--cmd", default="")
    ap.add_argument("--max-kernels", type=int, default=16)
    a = ap.parse_args()
    print(f"# {a.title}\n")
    if a.cmd:
        print(f"Command: `{a.cmd}`\n")
    if a.launches:
        print(f"## Launch list (`{a.launches}`)\n")
        print(launches_table(a.launches, a.first_kernel))
    if a.rep:
        print(f"## `ncu --set full` (`{a.rep}`), first launch of each kernel\n")
        print(rep_table(a.rep, a.max_kernels))
