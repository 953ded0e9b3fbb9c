#!/usr/bin/env python
"""Condense `ncu --page raw --csv` dumps (profiles/*_full_raw.csv) into one
JSON line per kernel launch: time, DRAM traffic, FMA-pipe use, instruction
count, registers, the five largest warp-stall reasons.

  python scripts/ncu_summary.py profiles/r02m_forward_expect_full_raw.csv ...
"""
import csv
import json
import sys

KEYS = {
    "time_us": "gpu__time_duration.sum",
    "dram_read_GB": "dram__bytes_read.sum",
    "dram_write_GB": "dram__bytes_write.sum",
    "dram_pct_of_peak": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_slots_pct": "sm__inst_executed.sum.pct_of_peak_sustained_elapsed",
    "warp_instructions": "smsp__inst_executed.sum",
    "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
}


def main(paths):
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        # pc-sampling counts of the stall reasons (shares of all samples)
        stall = [(h, i) for i, h in enumerate(hdr)
                 if h.startswith("smsp__pcsamp_warps_issue_stalled_") and
                 not h.endswith("_not_issued")]
        for r in data:
            out = {"file": path.split("/")[-1], "kernel": r[col["Kernel Name"]]}
            for k, name in KEYS.items():
                if name in col and r[col[name]] not in ("", "n/a"):
                    v = float(r[col[name]].replace(",", ""))
                    u = units[col[name]]
                    if k.endswith("_GB"):
                        v *= {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9,
                              "Tbyte": 1e3}.get(u, 1.0)
                    if k == "time_us":
                        v *= {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
                    out[k] = round(v, 4)
            st = []
            for h, i in stall:
                try:
                    st.append((float(r[i].replace(",", "")), h.split("issue_stalled_")[1]))
                except ValueError:
                    pass
            tot = sum(v for v, _ in st) or 1.0
            st.sort(reverse=True)
            out["top_stalls_pct_of_samples"] = {n: round(100.0 * v / tot, 1) for v, n in st[:6]}
            print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1:])
