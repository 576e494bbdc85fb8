"""Summarise an `ncu --set full` capture of tools/prof_pair.py into the small JSON bench.py reads for `roofline.traffic`.

    ncu --set full --clock-control none --import-source on -k regex:"k_tpanel|k_gather_panel|k_gather_u|k_gather_smem|k_tgather_t|k_tscatter_smem" \
        -c 12 -o gpurun_out/r2_pair python tools/prof_pair.py c3 1
    ncu -i gpurun_out/r2_pair.ncu-rep --page raw --csv > profiles/r2_pair_c3_ncu_raw.csv
    python tools/ncu_summary.py profiles/r2_pair_c3_ncu_raw.csv c3 4 > profiles/r2_root_pair_traffic.json

The LAST launch of every distinct kernel is taken (tools/prof_pair.py runs the pair reps + 2 times on the root node), the
dram__bytes_read.sum + dram__bytes_write.sum of those launches are added up: that is the DRAM traffic of one root-level pair.
"""
import csv, json, sys

path, workload, bits = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, key):
    v = float(r[ix[key]].replace(",", ""))
    u = units[ix[key]].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1,
             "nsecond": 1e-6, "second": 1e3, "s": 1e3}.get(u, 1)
    return v * scale


last = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    last[r[ix["Kernel Name"]]] = r
kernels = {}
for name, r in last.items():
    kernels[name] = {"ms": num(r, "gpu__time_duration.sum"), "dram_read_bytes": num(r, "dram__bytes_read.sum"),
                     "dram_write_bytes": num(r, "dram__bytes_write.sum"),
                     "issue_active_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                     "dram_pct": float(r[ix["dram__throughput.avg.pct_of_peak_sustained_elapsed"]]) if "dram__throughput.avg.pct_of_peak_sustained_elapsed" in ix else None}
out = {"workload": workload, "panel_bits": bits, "source_csv": path, "kernels": kernels,
       "pair_ms_under_ncu": sum(k["ms"] for k in kernels.values()),
       "pair_dram_bytes": sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in kernels.values())}
print(json.dumps(out, indent=1))
