#!/bin/bash
# Same-box A/B of two builds of libsdt_b200.so through bench.py (cfg2, CUDA-graph replay): alternates the libraries, one JSON
# line per run into gpurun_out/<tag>_<name>_<i>.json, then prints ms/step, the forward / dX fractions and the summed-source rows.
#   tools/ab_builds.sh <tag> <other.so> [rounds]
set -u
tag=$1; other=$2; rounds=${3:-2}
lib=scal_sdt_b200/_build/libsdt_b200.so
cp $lib /tmp/lib_new.so
mkdir -p gpurun_out
for i in $(seq 1 $rounds); do
  cp $other $lib; python bench.py --workload ${WORKLOAD:-cfg2} --no-cpu-baseline --no-torch-baseline > gpurun_out/${tag}_other_$i.json 2>/dev/null
  cp /tmp/lib_new.so $lib; python bench.py --workload ${WORKLOAD:-cfg2} --no-cpu-baseline --no-torch-baseline > gpurun_out/${tag}_new_$i.json 2>/dev/null
done
cp /tmp/lib_new.so $lib
python - "$tag" <<'PY'
import glob, json, sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d["roofline"]; dx = r["backward"]["dx_gemm"]
    rows = [(x["M"], x["contraction"], x["sources_per_launch"], round(x["avg_us"], 2)) for x in dx.get("by_shape", []) if x["sources_per_launch"] == 3 and x["writes_dx"]]
    print(f"{f}: {d['ms_per_step']:.3f} ms/step  fwd {r['frac']:.4f} ({r['forward_gemm_seconds_per_step']*1e3:.3f} ms)  dX {dx['frac']:.4f} ({dx['seconds_per_step']*1e3:.3f} ms)  step kernels {d.get('step_kernel_us')}  summed dX {rows}")
PY
