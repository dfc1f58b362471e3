#!/usr/bin/env python
"""Digest of `ncu --set full` captures (exported with `ncu -i X.ncu-rep --page raw --csv`) into the small JSON that
profiles/ keeps and bench.py reads (only for the configuration recorded in `_config`).
usage: python tools/ncu_digest.py out.json M_loc B k  raw1.csv [raw2.csv ...]"""
import csv
import json
import sys

KEYS = {
    "duration_us": ("gpu__time_duration.sum", 1.0),
    "sm_clock_ghz_during_capture": ("sm__cycles_elapsed.avg.per_second", 1.0),
    "dram_read_mbyte": ("dram__bytes_read.sum", 1.0),
    "dram_write_mbyte": ("dram__bytes_write.sum", 1.0),
    "dram_pct_of_peak": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "warp_instructions": ("smsp__inst_executed.sum", 1.0),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    "issue_active_pct_busiest_smsp": ("smsp__issue_active.max.pct_of_peak_sustained_active", 1.0),
    "alu_pipe_pct": ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
    "fma_pipe_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0),
    "xu_pipe_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
    "lsu_pipe_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1.0),
    "tensor_pipe_cycles_active_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    "registers": ("launch__registers_per_thread", 1.0),
    "block": ("launch__block_size", 1.0),
    "grid": ("launch__grid_size", 1.0),
}
STALLS = ["barrier", "branch_resolving", "dispatch_stall", "long_scoreboard", "math_pipe_throttle", "mio_throttle",
          "no_instruction", "not_selected", "short_scoreboard", "wait"]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


out = {"_config": {"M_loc": int(sys.argv[2]), "B": int(sys.argv[3]), "ks": [int(sys.argv[4])]},
       "_source": "ncu --set full --clock-control none --import-source on; bench.py --rows 20000 --steps 3 --warmup 3 "
                  "--no-cpu --no-e2e (tools/gpu_calls/r2_call17.sh); units as ncu prints them (Mbyte = 1e6 bytes)"}
for path in sys.argv[5:]:
    rows = list(csv.reader(open(path)))
    h = rows[0]
    for r in rows[2:]:
        d = dict(zip(h, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("nadm::", "")
        e = {k: num(d.get(m, "")) for k, (m, _) in KEYS.items()}
        if e["dram_read_mbyte"] is not None and e["dram_write_mbyte"] is not None:
            e["dram_bytes"] = (e["dram_read_mbyte"] + e["dram_write_mbyte"]) * 1e6
        e["stalls_per_issue"] = {s: num(d.get(f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio", ""))
                                 for s in STALLS}
        out[name] = e
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps({k: (v if k.startswith("_") else {kk: v[kk] for kk in ("duration_us", "dram_bytes", "issue_active_pct")})
                  for k, v in out.items()}, indent=1))
