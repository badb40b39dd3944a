import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tenncor_b200 import cabi
cabi.init(0)
peaks = json.load(open("MEASURED_PEAKS.json"))
for rep in range(2):
    r = bench.hbm_micro(cabi, peaks)
    print({k: (v["GBps"], v["frac"]) for k, v in r.items() if k != "_note" and ("dim1" in k or "SLICE" in k or "PAD" in k or "ARGMAX" in k or "bias" in k)}, flush=True)
