"""DRAM traffic of the step's GEMM launches from an ncu CSV log -> profiles/gemm_traffic.json (read by bench.py's
roofline.traffic).  The log comes from ONE eager step of the benchmark workload:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:gemm_sm100_kernel --launch-skip <2 warm-up steps> --launch-count <launches/step> --csv \
        --log-file gpurun_out/gemm_traffic.csv python bench.py --no-graph --no-roofline --no-cpu-baseline \
        --no-gpu-eager --steps 1 --warmup 3
    python tools/gemm_traffic.py gpurun_out/gemm_traffic.csv [--batch 32 --frames 8 --geom base]
"""
import argparse
import csv
import json
import os
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--geom", default="base")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "gemm_traffic.json"))
    a = ap.parse_args()
    rows = [r for r in csv.reader(l for l in open(a.csv) if l.startswith('"'))]
    hdr = rows[0]
    iid, ik, im, iu, iv = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
             "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
    per = defaultdict(dict)
    for r in rows[1:]:
        per[(r[iid], r[ik].split("(")[0])][r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    tot_b = tot_t = 0.0
    by_kernel = defaultdict(lambda: [0, 0.0, 0.0])
    for (lid, name), m in per.items():
        b = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        t = m.get("gpu__time_duration.sum", 0.0)
        tot_b += b
        tot_t += t
        k = by_kernel[name]
        k[0] += 1; k[1] += b; k[2] += t
    n = len(per)
    out = {"geom": a.geom, "batch": a.batch, "frames": a.frames, "launches": n, "bytes_per_launch": tot_b / max(n, 1),
           "total_bytes_per_step": tot_b, "ncu_time_s_per_step": tot_t,
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over every gemm_sm100_kernel launch of one eager step",
           "by_kernel": {k: {"launches": v[0], "bytes": v[1], "ncu_time_s": v[2]} for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1][1])}}
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("launches", "bytes_per_launch", "total_bytes_per_step", "ncu_time_s_per_step")}))


if __name__ == "__main__":
    main()
