"""Registers / stack (spill) / static shared memory per kernel family of libtcr_b200.so, fp32 instantiations (no GPU needed):
    cuobjdump -res-usage tenncor_b200/lib/libtcr_b200.so | python tools/res_usage.py > profiles/r1_res_usage.md"""
import collections
import re
import subprocess
import sys


def main():
    rows, name = [], None
    for line in sys.stdin:
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            rows.append((name,) + tuple(int(x) for x in m.groups()))
            name = None
    dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
    fam = collections.OrderedDict()
    for (_, reg, stack, shared, local), d in zip(rows, dem):
        d = d.replace("(anonymous namespace)::", "")
        base = re.sub(r"^void ", "", re.sub(r"[<(].*", "", d))
        f = fam.setdefault(base, {"n": 0, "reg": [], "stack": [], "shared": []})
        f["n"] += 1
        f["reg"].append(reg)
        f["stack"].append(stack)
        f["shared"].append(shared)
    print("# Resource usage per kernel family - libtcr_b200.so (sm_100a), `cuobjdump -res-usage` (tools/res_usage.py)\n")
    print("`stack` > 0 means local-memory frames (spills or indexed local arrays); dynamic shared memory (the GEMM rings) is not listed by cuobjdump.\n")
    print("| kernel family | instantiations | registers (min-max) | stack bytes (max) | static smem bytes (max) |")
    print("|---|---|---|---|---|")
    for base, f in sorted(fam.items()):
        print("| `%s` | %d | %d-%d | %d | %d |" % (base, f["n"], min(f["reg"]), max(f["reg"]), max(f["stack"]), max(f["shared"])))


if __name__ == "__main__":
    main()
