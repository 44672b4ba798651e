#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --page raw --csv` export of a --set full capture: per C-ABI entry
point, the mean DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and duration per launch.  bench.py
reads it for `roofline.traffic`.  Usage: tools/ncu_traffic.py X_raw.csv [profiles/ncu_traffic.json]"""
import csv
import json
import sys

ENTRY = {"linear_tma_kernel": "pfo_linear_tf32", "wgrad_tma_kernel": "pfo_wgrad_tf32",
         "attn_nbr_fwd_kernel": "pfo_attn_nbr_fwd", "attn_nbr_bwd_kernel": "pfo_attn_nbr_bwd",
         "mv_select_kernel": "pfo_mv_select", "neighbor_": "pfo_neighbor_sample", "bpr_kernel": "pfo_bpr",
         "store_messages_kernel": "pfo_store_messages", "persist_rank_kernel": "pfo_persist_rank",
         "gather_state_kernel": "pfo_gather_state", "cell_forward_kernel": "pfo_cell_forward",
         "cell_backward_kernel": "pfo_cell_backward", "mark_nodes_kernel": "pfo_mark_nodes"}


def main(path, out):
    rows = list(csv.reader(open(path)))
    head, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(head)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
    agg = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        ep = next((v for k, v in ENTRY.items() if k in name), None)
        if ep is None:
            continue
        f = lambda k: float(r[idx[k]].replace(",", ""))
        b = sum(f(k) * scale.get(units[idx[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = f("gpu__time_duration.sum") * tscale.get(units[idx["gpu__time_duration.sum"]], 1.0)
        a = agg.setdefault(ep, {"launches": 0, "bytes": 0.0, "us": 0.0})
        a["launches"] += 1; a["bytes"] += b; a["us"] += t
    res = {k: {"dram_bytes_per_launch": v["bytes"] / v["launches"], "ncu_us_per_launch": v["us"] / v["launches"],
               "launches_profiled": v["launches"], "source": "profiles/r2f_ncu_full_summary.txt (ncu --set full --clock-control none over the kernels of the C-ABI entry points, one eager training step after warm-up, bs 8192, final round-2 kernels)"} for k, v in agg.items()}
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    for k, v in sorted(res.items()):
        print(f"{k:24s} {v['dram_bytes_per_launch'] / 1e6:9.2f} MB/launch {v['ncu_us_per_launch']:8.1f} us x{v['launches_profiled']}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "profiles/ncu_traffic.json")
