"""Host-side cost of one asynchronous trainer step (ctypes + C++ host layer + CUDA API calls), measured with the GPU
NOT being the bottleneck: wall time per step_async call while at most 6 steps are outstanding."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from taper_b200 import host

host.set_device(0)
for batch in (64, 512):
    m = host.Model(host.MLP_784_128_10, 0)
    tr = host.Trainer(m, "adam", lr=1e-3)
    pins = [(host.PinnedArray((batch, 784)), host.PinnedArray((batch,))) for _ in range(12)]
    rng = np.random.default_rng(0)
    for px, py in pins:
        px.array[...] = rng.random((batch, 784), dtype=np.float32); py.array[...] = rng.integers(0, 10, batch)
    for j in range(20):
        tr.step_async(pins[j % 12][0].array, pins[j % 12][1].array, pinned=True); tr.fetch()
    n = 20000
    t_call = 0.0
    t0 = time.perf_counter()
    for i in range(n):
        if tr.pending() >= 6:
            tr.fetch()
        a = time.perf_counter()
        tr.step_async(pins[i % 12][0].array, pins[i % 12][1].array, pinned=True)
        t_call += time.perf_counter() - a
    while tr.pending():
        tr.fetch()
    host.sync()
    t1 = time.perf_counter()
    print(f"batch {batch}: {1e6 * (t1 - t0) / n:.1f} us/step wall, {1e6 * t_call / n:.1f} us inside step_async (host)")
