"""Summarise which kernels of libtcr_b200.so contain tensor-core / TMEM / TMA SASS (no GPU needed):
    cuobjdump -sass tenncor_b200/lib/libtcr_b200.so | python tools/sass_evidence.py > profiles/r1_sass_evidence.md
Mnemonics per /opt/skills/guides/B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG / UBLKCP, mma.sync -> HMMA."""
import collections
import re
import subprocess
import sys

PATTERNS = collections.OrderedDict([
    ("mma", re.compile(r"\bUTC\w*MMA")), ("ldtm", re.compile(r"\bLDTM")), ("tma", re.compile(r"\bUTMALDG")),
    ("tma_store", re.compile(r"\bUTMASTG|\bUBLKCP")), ("hmma", re.compile(r"\bHMMA")), ("syncs", re.compile(r"\bSYNCS"))])


def main():
    counts, cur = collections.OrderedDict(), None
    for line in sys.stdin:
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        if cur is not None:
            for key, pat in PATTERNS.items():
                if pat.search(line):
                    cur[key] += 1
    names = list(counts)
    demangled = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    short = {n: re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "")) for n, d in zip(names, demangled)}
    print("# SASS evidence - tenncor_b200/lib/libtcr_b200.so (sm_100a), `cuobjdump -sass` in the build container (tools/sass_evidence.py)\n")
    print("Kernels whose SASS contains tensor-core / TMEM / TMA instructions (mnemonic table: B200_PROFILING.md, \"What proves a Blackwell-native kernel\").")
    print("HMMA (legacy mma.sync / wmma) anywhere in the library: %s.\n" % ("none" if not any(c["hmma"] for c in counts.values()) else "PRESENT"))
    print("| kernel | UTC*MMA (tcgen05.mma) | LDTM (tcgen05.ld) | UTMALDG (TMA load) | SYNCS (mbarrier) |")
    print("|---|---|---|---|---|")
    n = 0
    for name, c in counts.items():
        if c["mma"] or c["tma"] or c["ldtm"]:
            print("| `%s` | %d | %d | %d | %d |" % (short[name], c["mma"], c["ldtm"], c["tma"], c["syncs"]))
            n += 1
    print("\n%d of %d kernels; the others are the HBM-bound elementwise / reduce / layout / skinny-product / patch kernels, which by design stay off the tensor pipe." % (n, len(counts)))


if __name__ == "__main__":
    main()
