"""Condense an `ncu --page raw --csv` export into one JSON record per kernel (averaged over its launches):
duration, DRAM bytes read / written, DRAM and SM throughput, tensor-pipe activity, occupancy, registers.

    python tools/ncu_summary.py gpurun_out/r2_ncu_full_iter_raw.csv profiles/r2_ncu_full_iter.json
"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src, newline="")))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[hdr_i], rows[hdr_i + 1]
col = {name: i for i, name in enumerate(hdr)}


def find(*subs):
    for name in hdr:
        if all(s in name for s in subs):
            return name
    return None


WANT = OrderedDict([
    ("time_us", find("gpu__time_duration.sum")),
    ("dram_read_bytes", find("dram__bytes_read.sum") and "dram__bytes_read.sum"),
    ("dram_write_bytes", find("dram__bytes_write.sum") and "dram__bytes_write.sum"),
    ("dram_throughput_pct", find("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
    ("sm_throughput_pct", find("sm__throughput.avg.pct_of_peak_sustained_elapsed")),
    ("tensor_pipe_active_pct", find("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")),
    ("issue_active_pct", find("smsp__issue_active.avg.pct_of_peak_sustained_active") or find("sm__inst_executed", "pct")),
    ("achieved_occupancy_pct", find("sm__warps_active.avg.pct_of_peak_sustained_active")),
    ("registers_per_thread", find("launch__registers_per_thread")),
    ("grid_size", find("launch__grid_size")),
    ("block_size", find("launch__block_size")),
])
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(row, name):
    if not name or name not in col:
        return None
    try:
        v = float(row[col[name]].replace(",", ""))
    except ValueError:
        return None
    return v * SCALE.get(units[col[name]], 1.0)


agg = defaultdict(lambda: defaultdict(list))
for r in rows[hdr_i + 2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).strip()
    name = re.sub(r"^void\s+|\(anonymous namespace\)::", "", name)
    for k, c in WANT.items():
        v = val(r, c)
        if v is not None:
            agg[name][k].append(v)
out = []
for name, d in agg.items():
    rec = {"kernel": name, "launches": len(d.get("time_us", []))}
    for k, vs in d.items():
        rec[k] = round(sum(vs) / len(vs), 3)
    out.append(rec)
out.sort(key=lambda r: -r.get("time_us", 0) * r["launches"])
json.dump({"source": src, "columns": {k: v for k, v in WANT.items()}, "kernels": out}, open(dst, "w"), indent=1)
for r in out:
    print(f'{r["kernel"][:48]:48s} n={r["launches"]:3d} {r.get("time_us", 0):9.1f} us  dram {r.get("dram_read_bytes", 0) / 1e6:8.2f}+{r.get("dram_write_bytes", 0) / 1e6:8.2f} MB'
          f'  dram% {r.get("dram_throughput_pct", 0):5.1f}  tensor% {r.get("tensor_pipe_active_pct", 0):5.1f}  sm% {r.get("sm_throughput_pct", 0):5.1f}')
