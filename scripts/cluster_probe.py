import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taper_b200
from taper_b200 import capi
ctx = taper_b200.Ctx(0)
for bn in (16, 32, 64, 128):
    print(bn, {c: capi.lib.tpdbg_max_clusters(bn, c) for c in (1, 2, 4, 8)})
