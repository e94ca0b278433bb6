"""Per-launch DRAM traffic of the backward tile pass from an ncu launch list (`ncu --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv ...`), written as the small JSON bench.py reads for
`roofline.traffic`:   python scripts/ncu_traffic.py gpurun_out/launches.csv 30 profiles/r2_traffic_bwd_n30.json"""
import csv
import hashlib
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
per = {}
for r in rows:
    per.setdefault(int(r[0]), {"name": r[4]})[r[12]] = float(r[14].replace(",", ""))
import re  # noqa: E402
bwd = [v for v in per.values() if re.search(r"k_tile12(_gs?)?(<|ILi)2", v["name"])]   # static, axis-aware and split-barrier passes
fwd = [v for v in per.values() if re.search(r"k_tile12(_gs?)?(<|ILi)1", v["name"])]
h = hashlib.sha1()
for f in sorted(glob.glob(os.path.join(ROOT, "qradient_b200", "csrc", "*"))):
    h.update(open(f, "rb").read())


def mean(vs, key):
    return sum(v.get(key, 0.0) for v in vs) / max(len(vs), 1)


out = {"n_qubits": int(sys.argv[2]), "source_hash": h.hexdigest()[:12], "launch_list": os.path.basename(sys.argv[1]),
       "backward_launches": len(bwd), "dram_bytes_per_launch": mean(bwd, "dram__bytes_read.sum") + mean(bwd, "dram__bytes_write.sum"),
       "dram_read_per_launch": mean(bwd, "dram__bytes_read.sum"), "dram_write_per_launch": mean(bwd, "dram__bytes_write.sum"),
       "ns_per_launch_under_ncu": mean(bwd, "gpu__time_duration.sum"), "algorithmic_bytes_per_launch": 64.0 * 2.0 ** int(sys.argv[2]),
       "forward_launches": len(fwd), "forward_dram_bytes_per_launch": mean(fwd, "dram__bytes_read.sum") + mean(fwd, "dram__bytes_write.sum")}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out))
