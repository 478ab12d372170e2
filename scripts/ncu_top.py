"""Summarise an ncu report: per-kernel duration + top stall sites.  usage: ncu_top.py report.ncu-rep [max_kernels]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
maxk = int(sys.argv[2]) if len(sys.argv) > 2 else 4
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
def col(name): return hdr.index(name) if name in hdr else None
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]
for r in rows[2:2 + maxk]:
    print({w: r[col(w)] for w in want if col(w) is not None})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
k = 0; hdr2 = None; cur = []
blocks = []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        if cur: blocks.append((name, hdr2, cur))
        name = r[1]; cur = []; continue
    if r and r[0] == "Address":
        hdr2 = r; continue
    if hdr2 and len(r) == len(hdr2): cur.append(r)
if cur: blocks.append((name, hdr2, cur))
for name, h, out in blocks[:maxk]:
    si, so, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    tot = sum(int(r[si]) for r in out)
    print("\n==", name[:100], "samples", tot)
    for r in sorted(out, key=lambda r: -int(r[si]))[:14]:
        st = {x[6:]: r[i] for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x and r[i] not in ("0", "")}
        print(f"{int(r[si]):6d} {r[ie]:>8s}  {r[so][:60]:60s} {st}")
