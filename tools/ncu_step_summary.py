"""Aggregate an `ncu --csv` metric log of ONE training step into a per-kernel JSON summary (runs without a GPU).

    ncu --profile-from-start off --clock-control none \
        -k "regex:lora_|gn_|ln_|geglu|residual_bias|adamw|ema_|mse_loss|noise_target" \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
        --csv --log-file gpurun_out/r02_sdt_metrics.csv python bench.py --profile-step
    python tools/ncu_step_summary.py gpurun_out/r02_sdt_metrics.csv > profiles/r02_kernels_per_step_ncu.json

Per kernel name (template arguments kept, signature dropped): launches, total / average device time (cold cache, serialised: compare
SHARES), DRAM bytes read / written, bytes per launch, time-weighted tensor-pipe activity.  bench.py reads `dram_*_MB` of the
lora_gemm* entries for `roofline.traffic`.
"""
import collections
import csv
import json
import re
import sys


def main(path):
    rows = collections.defaultdict(dict)
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for rec in csv.DictReader(lines):
        rows[(rec["ID"], rec["Kernel Name"])][rec["Metric Name"]] = (float(rec["Metric Value"].replace(",", "")), rec["Metric Unit"])
    agg = collections.OrderedDict()
    for (_id, name), m in rows.items():
        short = re.sub(r"^void\s+", "", name)
        short = re.sub(r"\(.*$", "", short).replace("sdt::", "")
        a = agg.setdefault(short, {"launches": 0, "total_us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0, "_tp": 0.0})
        dur, unit = m.get("gpu__time_duration.sum", (0.0, "ns"))
        us = dur / 1e3 if unit in ("ns", "nsecond") else (dur if unit in ("us", "usecond") else dur * 1e3)
        a["launches"] += 1
        a["total_us"] += us
        a["dram_read_MB"] += m.get("dram__bytes_read.sum", (0.0, ""))[0] / 1e6 * _scale(m.get("dram__bytes_read.sum", (0.0, "byte"))[1])
        a["dram_write_MB"] += m.get("dram__bytes_write.sum", (0.0, ""))[0] / 1e6 * _scale(m.get("dram__bytes_write.sum", (0.0, "byte"))[1])
        tp = [v for k, v in m.items() if k.startswith("sm__pipe_tensor")]
        if tp:
            a["_tp"] += tp[0][0] * us
    out = collections.OrderedDict()
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["total_us"]):
        n = a["launches"]
        out[name] = {"launches": n, "total_us": round(a["total_us"], 1), "avg_us": round(a["total_us"] / n, 2),
                     "dram_read_MB": round(a["dram_read_MB"], 1), "dram_write_MB": round(a["dram_write_MB"], 1),
                     "dram_bytes_per_launch": int((a["dram_read_MB"] + a["dram_write_MB"]) * 1e6 / n),
                     "time_weighted_tensor_pipe_pct": round(a["_tp"] / a["total_us"], 1) if a["total_us"] else 0.0,
                     "dram_GBps": round((a["dram_read_MB"] + a["dram_write_MB"]) * 1e6 / (a["total_us"] * 1e-6) / 1e9, 1) if a["total_us"] else 0.0}
    json.dump(out, sys.stdout, indent=1)
    print()


def _scale(unit):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


if __name__ == "__main__":
    main(sys.argv[1])
