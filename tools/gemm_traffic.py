#!/usr/bin/env python
"""profiles/gemm_dram_traffic.json from the ncu CSVs written by tools/ncu_capture.sh step 2 (per-launch DRAM bytes of the
GEMM kernels of one frame).  bench.py reads the mean as roofline.traffic."""
import csv
import json
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = {}
    for r in csv.DictReader(lines):
        d = per.setdefault(r["ID"], dict(name=r["Kernel Name"][:48], rd=0.0, wr=0.0, ns=0.0))
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(u, 1)
        if "bytes_read" in r["Metric Name"]:
            d["rd"] = v * mult
        elif "bytes_write" in r["Metric Name"]:
            d["wr"] = v * mult
        else:
            d["ns"] = v * mult
    return list(per.values())


def main():
    tag = sys.argv[1]
    out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                     "over every conv/fc GEMM launch of one frame (tools/ncu_capture.sh %s, step 2)" % tag}
    for v in (3, 2):
        rows = load("gpurun_out/%s_gemm_dram_v%d.csv" % (tag, v))
        tot = sum(r["rd"] + r["wr"] for r in rows)
        out["views%d" % v] = dict(launches=len(rows), total_dram_bytes=tot, mean_dram_bytes_per_launch=tot / max(1, len(rows)),
                                  total_ns_under_ncu=sum(r["ns"] for r in rows),
                                  per_launch=[dict(kernel=r["name"], dram_MB=(r["rd"] + r["wr"]) / 1e6, us=r["ns"] / 1e3) for r in rows])
    json.dump(out, open("profiles/gemm_dram_traffic.json", "w"), indent=1)
    print({k: (v["launches"], v["mean_dram_bytes_per_launch"]) for k, v in out.items() if k.startswith("views")})


if __name__ == "__main__":
    main()
