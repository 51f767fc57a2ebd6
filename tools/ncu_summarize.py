"""Reduces an `ncu --set full` capture of ONE plain forward (tools/ncu_forward.py) to the committed evidence:

    python tools/ncu_summarize.py <capture.ncu-rep> <optable.json> <out_summary.csv> <out_traffic.json> [batch]

* summary CSV: one row per kernel launch, named after the op of bench.py's per-op table it implements (launch order =
  op order; ops fused into a neighbour - "(fused)" in the table - launch nothing), with duration, DRAM bytes, unit
  throughputs, issue-slot utilisation, registers, dynamic shared memory, executed instructions;
* traffic JSON: dram__bytes_read.sum + dram__bytes_write.sum per launch keyed by op name (`roofline.traffic` of bench.py).
Runs where ncu is installed (no GPU needed to read a report)."""
import csv
import json
import subprocess
import sys

rep, optable, out_csv, out_json = sys.argv[1:5]
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 32
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ops = [o["op"] for o in json.load(open(optable))["ops"] if "(fused)" not in o["op"]]
if len(data) != len(ops):
  print("warning: %d launches in the capture, %d kernel-launching ops in the table: matching the LAST %d launches" %
        (len(data), len(ops), min(len(data), len(ops))))
  data = data[-len(ops):]
  ops = ops[-len(data):]
cols = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
idx = [hdr.index(c) for c in cols if c in hdr]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
with open(out_csv, "w", newline="") as f:
  w = csv.writer(f)
  w.writerow(["op"] + ["%s [%s]" % (hdr[i], units[i]) for i in idx])
  for op, r in zip(ops, data):
    w.writerow([op] + [r[i] for i in idx])
traffic = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full --clock-control none` "
                    "capture of a plain forward at batch %d (%s); key = op name of bench.py's per-op table" % (batch, out_csv)}
ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
for op, r in zip(ops, data):
  traffic[op] = {"dram_read_bytes": float(r[ir]) * scale.get(units[ir], 1.0), "dram_write_bytes": float(r[iw]) * scale.get(units[iw], 1.0),
                 "batch": batch, "ncu_duration_%s" % units[it]: float(r[it]), "capture": out_csv.split("/")[-1]}
json.dump(traffic, open(out_json, "w"), indent=1)
print("wrote", out_csv, out_json, len(ops), "ops")
