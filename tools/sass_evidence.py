"""Opcode evidence for the shipped library (runs on the build box, no GPU): disassembles libmgv.so with cuobjdump and
counts, per kernel, the SASS mnemonics that identify the Blackwell paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
LDTM = tcgen05.ld, UTMALDG = TMA load, SYNCS = mbarrier, HMMA = legacy mma.sync, plus registers / shared memory from
--dump-resource-usage.  Writes profiles/r2_sass_opcodes.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "melspec_gpt_vqvae_b200", "libmgv.so")
KEYS = ["UTCHMMA", "UTCMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMAPF", "UTMASTG", "SYNCS", "HMMA", "FFMA", "FFMA2", "LDGSTS", "REDG", "ATOMG", "MUFU"]
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", SO], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
rows = []
cur, cnt, total = None, None, 0
def flush():
    if cur:
        rows.append((cur, dict(cnt), total))
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        flush()
        cur, cnt, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        total += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or (k in ("UTCHMMA", "UTCMMA", "UTCQMMA") and op.startswith(k)):
                cnt[k] += 1
flush()
out = ["# SASS opcode counts per kernel of melspec_gpt_vqvae_b200/libmgv.so (tools/sass_evidence.py; cuobjdump -sass, sm_100a)",
       "# UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, SYNCS = mbarrier ops, HMMA = legacy mma.sync",
       "%-78s %5s %5s %6s %7s  %s" % ("kernel", "regs", "stack", "smem_s", "instrs", "opcodes")]
tot = collections.Counter()
for name, c, n in sorted(rows, key=lambda r: demangle(r[0])):
    d = demangle(name)
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*$", "", d)[:78]
    u = usage.get(name, (0, 0, 0))
    ops = " ".join("%s=%d" % (k, c[k]) for k in KEYS if c.get(k))
    out.append("%-78s %5d %5d %6d %7d  %s" % (d, u[0], u[1], u[2], n, ops))
    tot.update(c)
out.append("")
out.append("TOTAL " + " ".join("%s=%d" % (k, tot[k]) for k in KEYS if tot.get(k)))
path = os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[-1:]))
print("wrote", path, "(%d kernels)" % len(rows))
