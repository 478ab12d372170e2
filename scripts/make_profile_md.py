"""Regenerates the tables of profiles/r1_device_tape.md from the raw files under profiles/ (bench lines, ncu launch lists)."""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
md = open(os.path.join(P, "r1_device_tape.md")).read()
bench = {c: json.loads(open(os.path.join(P, f"r1_bench_{c}.json")).read()) for c in ("cfg2", "cfg1", "example_mlp", "cfg4", "cnn2", "cnn5")}
rows = "\n".join(f"| {c} | {bench[c]['config']['step_path'][:46]} | {bench[c]['value']:.3e} | {bench[c]['ms_per_step']*1e3:.1f} | "
                 f"{bench[c]['launches_per_step']:.0f} | {bench[c]['e2e']['value']:.3e} | {bench[c]['cpu_baseline']['value']:.3e} |" for c in bench)
md = re.sub(r"(\| config \| step path .*?\n\|---.*?\n)(?:\|.*\n)+", lambda m: m.group(1) + rows + "\n", md, count=1)
lr = [r for r in csv.reader(open(os.path.join(P, "r1_launches_cnn5.csv"))) if len(r) > 10 and r[0].isdigit()]
idx = [i for i, r in enumerate(lr) if "gather_batch" in r[4]]
step = lr[idx[-2]:idx[-1]]
tot = sum(float(r[-1]) for r in step)
agg = {}
for r in step:
    name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += float(r[-1]) / 1e3
lines = "\n".join(f"| `{k}` | {n} | {t:.1f} | {100*t*1e3/tot:.1f}% |" for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]))
md = re.sub(r"\d+ launches, [\d.]+ µs summed \(cold-cache, serialised under ncu; `bench.py`: [\d.]+ µs per step\)",
            f"{len(step)} launches, {tot/1e3:.1f} µs summed (cold-cache, serialised under ncu; `bench.py`: {bench['cnn5']['ms_per_step']*1e3:.1f} µs per step)", md)
md = re.sub(r"(\| kernel \| launches \| µs \| share \|\n\|---\|---\|---\|---\|\n)(?:\|.*\n)+", lambda m: m.group(1) + lines + "\n", md, count=1)
open(os.path.join(P, "r1_device_tape.md"), "w").write(md)
print("ok")
