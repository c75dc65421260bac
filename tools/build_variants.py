"""Build A/B variants of libsplatter360.so into gpurun_variants/ (they travel to the GPU box; S360_LIB selects one).
    python tools/build_variants.py name=DEFINE[,DEFINE...] ...      e.g.  counters=S360_COUNTERS=1 qprime=S360_BWD_QPRIME=1"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splatter360_b200.csrc import build as B
os.makedirs(os.path.join(ROOT, "gpurun_variants"), exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    out = os.path.join(ROOT, "gpurun_variants", f"lib_{name}.so")
    B.build(force=True, out=out, defines=[d for d in defs.split(",") if d])
    print(out)
