"""cuobjdump -sass opcode counts per kernel of libaffgw.so -> profiles/<round>_sass_opcodes.txt (runs in the build container).

    python scripts/sass_opcodes.py [r02]
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
LIB = os.path.join(ROOT, "affganwriting_b200", "csrc", "libaffgw.so")
OPS = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "LDGSTS", "R2UR", "ELECT", "HMMA"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
per, cur = collections.OrderedDict(), None
total = collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|void ", "", name).split("(")[0]
        cur = per.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1).split(".")[0]
        cur[op] += 1
        total[op] += 1
with open(os.path.join(ROOT, "profiles", f"{R}_sass_opcodes.txt"), "w") as f:
    f.write(f"# cuobjdump -sass affganwriting_b200/csrc/libaffgw.so (sm_100a, {R}), opcode counts per kernel function (scripts/sass_opcodes.py)\n")
    f.write("# UTCHMMA = tcgen05.mma kind::f16 ; LDTM = tcgen05.ld ; UTMALDG = cp.async.bulk.tensor (tiled TMA) ; UBLKCP = cp.async.bulk ;\n")
    f.write("# LDGSTS = cp.async ; R2UR = vector -> uniform register move (0 per issued MMA since the issue loops run on the uniform\n")
    f.write("# datapath, 15 before) ; HMMA = legacy mma.sync (none expected)\n")
    f.write("# whole library: " + ", ".join(f"{o} {total[o]}" for o in OPS + ["UTCBAR", "SYNCS", "FFMA"]) + "\n")
    f.write("%-64s" % "kernel" + "".join("%9s" % o for o in OPS) + "\n")
    for name, c in sorted(per.items()):
        if c["UTCHMMA"] or c["HMMA"]:
            f.write("%-64s" % name[:64] + "".join("%9d" % c[o] for o in OPS) + "\n")
print({o: total[o] for o in OPS})
