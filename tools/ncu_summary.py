#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a markdown table: per launch duration, DRAM bytes, DRAM %, tensor-pipe %,
registers, achieved occupancy.   python tools/ncu_summary.py <rep> [<rep> ...] > profiles/xxx.md"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    m = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}
    return v * m.get(u, 1)


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print("## %s\n" % rep)
        print("| # | kernel | grid | regs | dur us | dram rd MB | dram wr MB | dram % | L2 MB | tensor % | sm % | warps % |")
        print("|---|---|---|---|---|---|---|---|---|---|---|---|")
        for n, r in enumerate(rows[2:]):
            def g(k, conv=True):
                i = idx.get(k)
                if i is None or r[i] == "":
                    return float("nan")
                return to_bytes(r[i], units[i]) if conv else float(r[i].replace(",", ""))
            name = r[idx["Kernel Name"]][:60]
            print("| %d | %s | %s | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |" % (
                n, name, r[idx["launch__grid_size"]], g("launch__registers_per_thread", False), g("gpu__time_duration.sum"),
                g("dram__bytes_read.sum") / 1e6, g("dram__bytes_write.sum") / 1e6,
                g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False), g("lts__t_bytes.sum") / 1e6,
                g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
                g("sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
                g("sm__warps_active.avg.pct_of_peak_sustained_active", False)))
        print()


if __name__ == "__main__":
    main()
