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
    steps = tc.profile_plan(5)
    total = sum(s["ms"] for s in steps)
    print("plan: %s" % tc.plan_stats())
    print("%-48s %-28s %9s %6s %9s" % ("step", "shape", "us", "%", "GB/s"))
    for s in steps:
        print("%-48s %-28s %9.1f %6.1f %9.1f" % (s["what"][:48], s["shape"][:28], s["ms"] * 1e3, 100 * s["ms"] / total, s["bytes"] / s["ms"] / 1e6))
    print("sum of steps: %.1f us" % (total * 1e3))
    if args.json:
        json.dump({"workload": args.workload, "steps": steps, "sum_ms": total}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
