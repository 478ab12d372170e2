"""Print the last complete step of an ncu launch list (--metrics gpu__time_duration.sum --csv): usage launch_list.py file.csv [first-kernel-substring]"""
import csv, sys
p = sys.argv[1]
first = sys.argv[2] if len(sys.argv) > 2 else "gather_batch"
rows = [r for r in csv.reader(open(p)) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
data = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hdr + 1:] if r[0].isdigit()]
idx = [i for i, (n, _) in enumerate(data) if first in n]
seg = data[idx[-2]:idx[-1]] if len(idx) >= 2 else data
tot = 0.0
for n, t in seg:
    print(f"{t / 1000:9.2f} us  {n[:130]}"); tot += t
print(f"{tot / 1000:9.2f} us  total, {len(seg)} launches")
