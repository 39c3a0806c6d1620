"""Per-kernel counts of the Blackwell-native SASS instructions in libstz.so (cuobjdump -sass): the proof that the hot kernels
are tcgen05 / TMEM / TMA code (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG / UTMASTG /
UTMAREDG, tcgen05.commit -> UTCBAR, legacy mma.sync -> HMMA).   python tools/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "styletts-zs_b200", "csrc", "libstz.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "UBLKCP", "SYNCS", "HMMA", "FFMA2", "LDGSTS"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# {os.path.relpath(so, ROOT)}: SASS instruction counts per kernel (cuobjdump -sass, sm_100a)")
print(f"{'kernel':90s} {'instr':>6s} " + " ".join(f"{k:>8s}" for k in KEYS))
for fn, c in counts.items():
    name = re.sub(r"\(.*", "", demangle(fn)).replace("void ", "").replace("stz::", "")
    print(f"{name[:90]:90s} {total[fn]:6d} " + " ".join(f"{c.get(k, 0):8d}" for k in KEYS))
