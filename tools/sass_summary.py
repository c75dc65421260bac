"""Per-kernel SASS summary of libsplatter360.so for profiles/: registers, shared memory, local-memory traffic (spills) and
the counts of the instructions that show which hardware path a kernel uses (TMA bulk copies UBLKCP, async copies LDGSTS,
two-wide FP32 FFMA2/FMUL2/FADD2, fire-and-forget reductions REDG, warp match / redux, local loads/stores LDL/STL).
    python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "splatter360_b200", "libsplatter360.so")
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(x) for x in m.groups())
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = collections.defaultdict(collections.Counter)
arch = None
cur = None
for line in sass.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        ops[cur][m.group(1)] += 1
KEYS = ["UBLKCP", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "REDG", "ATOMG", "ATOMS", "MATCH", "REDUX", "SHFL", "VOTE", "MUFU", "LDL", "STL"]
dem = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0].replace("s360::", "").replace("void ", "")
print(f"# {os.path.relpath(lib, ROOT)}: arch {arch}; per kernel: registers, static shared bytes, stack bytes (spills), SASS length, opcode counts")
print(f"{'kernel':58s} {'regs':>4s} {'smem':>6s} {'stack':>5s} {'sass':>5s} " + " ".join(f"{k:>6s}" for k in KEYS))
for fn in sorted(ops, key=lambda f: dem(f)):
    reg, stack, shared, local = usage.get(fn, (0, 0, 0, 0))
    c = ops[fn]
    print(f"{dem(fn)[:58]:58s} {reg:4d} {shared:6d} {stack:5d} {sum(c.values()):5d} " + " ".join(f"{c.get(k, 0):6d}" for k in KEYS))
tot = collections.Counter()
for c in ops.values():
    tot.update(c)
print("# library totals: " + " ".join(f"{k}={tot.get(k, 0)}" for k in KEYS))
