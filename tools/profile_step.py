"""Per-step device timings of one training step (planned evaluator): which kernels the step is made of,
their share of the step and achieved GB/s on algorithmic bytes. Usage: python tools/profile_step.py --workload c3"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--precision", default="3xtf32")
    ap.add_argument("--json", default="")
    ap.add_argument("--aggregate", action="store_true", help="group identical (step, shape) rows: for plans with thousands of steps")
    ap.add_argument("--repeats", type=int, default=5)
    args = ap.parse_args()
    import tenncor_b200 as tc
    tc.set_evaluator("plan")
    tc.set_matmul_precision(args.precision)
    cfg, gen, w = bench.build_config(args.workload)
    rng = np.random.default_rng(0)
    for f, arr in zip(cfg.feeds.values(), gen(rng)):
        f.assign(arr)
    for _ in range(3):
        cfg.train.calc()
        for f in cfg.feeds.values():
            f.touch()
    tc.sync()
    import re
    steps = tc.profile_plan(args.repeats)
    for s in steps:  # compact shapes: drop trailing unit ranks
        s["what"] = re.sub(r"(\\1)+\]", "]", s["what"])
        s["shape"] = re.sub(r"(\\1)+\]", "]", s["shape"])
    total = sum(s["ms"] for s in steps)
    print("plan: %s" % tc.plan_stats())
    print("%-48s %-28s %9s %6s %9s" % ("step", "shape", "us", "%", "GB/s"))
    if args.aggregate:
        groups = {}
        for s in steps:
            g = groups.setdefault((s["what"], s["shape"]), {"n": 0, "ms": 0.0, "bytes": 0})
            g["n"] += 1
            g["ms"] += s["ms"]
            g["bytes"] += s["bytes"]
        print("%-5s %-44s %-24s %10s %6s %8s %9s" % ("count", "step", "shape", "total us", "%", "us each", "GB/s"))
        for (what, shape), g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
            print("%-5d %-44s %-24s %10.1f %6.1f %8.2f %9.1f" % (g["n"], what[:44], shape[:24], g["ms"] * 1e3, 100 * g["ms"] / total,
                                                              g["ms"] * 1e3 / g["n"], g["bytes"] / g["ms"] / 1e6))
    else:
        for s in steps:
            print("%-48s %-28s %9.1f %6.1f %9.1f" % (s["what"][:48], s["shape"][:28], s["ms"] * 1e3, 100 * s["ms"] / total, s["bytes"] / s["ms"] / 1e6))
    print("sum of steps: %.1f us" % (total * 1e3))
    if args.json:
        json.dump({"workload": args.workload, "steps": steps, "sum_ms": total}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
