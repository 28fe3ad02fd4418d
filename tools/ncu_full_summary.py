"""Summarise an `ncu --set full` report (.ncu-rep) into a small JSON: per profiled launch the metrics the roofline discussion uses
(runs without a GPU: `ncu -i report --page raw --csv`).  `python tools/ncu_full_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.json`"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_of_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_KB",
    "sm__cycles_elapsed.max": "cycles_elapsed",
    "smsp__cycles_active.avg": "smsp_cycles_active",
}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.DictReader(io.StringIO(raw)))
    units, launches = rows[0], rows[1:]
    out = {"report": path.split("/")[-1], "how": "ncu --set full --clock-control none --import-source on (cold cache, serialised replays)",
           "launches": []}
    for r in launches:
        rec = {"kernel": r["Kernel Name"].split("(")[0].replace("void ", "")}
        for k, name in WANT.items():
            if k in r and r[k] != "":
                v = float(r[k].replace(",", ""))
                u = units.get(k, "")
                if u == "Gbyte":
                    v *= 1e3
                elif u == "Kbyte":
                    v /= 1e3
                elif u == "byte":
                    v /= 1e6 if name.endswith("_MB") else (1e3 if name.endswith("_KB") else 1)
                elif u in ("ns", "nsecond"):
                    v /= 1e3
                rec[name] = round(v, 3)
        out["launches"].append(rec)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1])
