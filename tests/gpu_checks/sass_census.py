"""Which kernels of libtpc.so contain which Blackwell instructions (cuobjdump -sass, no GPU needed).
python tests/gpu_checks/sass_census.py > profiles/r02f_sass_census.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = os.path.join(ROOT, "transformer-pointer-critic_b200", "libtpc.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "ELECT", "R2UR.BROADCAST", "F2FP.SATFINITE.F16",
        "HMMA.16816", "HMMA.1688", "MUFU.EX2", "MUFU.TANH", "RED.E.ADD", "SYNCS"]
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for c in cols:
            if op.startswith(c):
                counts[cur][c] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS census of libtpc.so (cuobjdump -sass, sm_100a): which kernels use which Blackwell instructions")
print("# UTCHMMA = tcgen05.mma (kind::f16 in the split-mode forward paths, kind::tf32 in the single-pass ones), LDTM / STTM = tcgen05.ld / st,")
print("# UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit, ELECT = elect.sync (MMA issue in uniform control flow),")
print("# R2UR.BROADCAST = the per-instruction operand broadcast ptxas emits around TMA / MMA instructions issued under a divergent predicate")
print("# (round 1: 305 in gemm_tc_kernel, all MMAs included), HMMA.* = legacy warp-level mma.sync (attention kernels only)\n")
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
for (k, c), d in zip(counts.items(), dem):
    if not any(c[x] for x in cols):
        continue
    name = re.sub(r"\(.*$", "", d.replace("(anonymous namespace)::", "")).replace("void ", "")
    print(f"| `{name}` | " + " | ".join(str(c[x]) for x in cols) + " |")
