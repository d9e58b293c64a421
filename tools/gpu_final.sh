#!/bin/bash
# GPU-box script: the round's final artifacts (writes gpurun_out/final_*)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -2 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "bench reference rc=$?"
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench default rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
for dt in tf32 bf16; do
  ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_$dt.csv \
      python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu --no-secondary --profile-step --dtype $dt > gpurun_out/final_ncu_$dt.log 2>&1
  echo "ncu $dt rc=$?"; wc -l gpurun_out/final_launches_$dt.csv
done
python - <<EOF
import json
for f in ("gpurun_out/final_bench_reference.json", "gpurun_out/final_bench_default.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, round(d["value"], 2), round(d["ms_per_step"], 3), d.get("e2e", {}).get("value"), d.get("clocks"), d.get("gpu_launches"))
            r = d.get("roofline")
            if r:
                print(" ", r["kernel"][:24], round(r["frac"], 3), r.get("traffic"), {k: round(v["frac"], 3) for k, v in r["other_kernels"].items()})
            for s in d.get("secondary", []):
                print("  secondary:", s.get("config", s)[:70] if isinstance(s.get("config", ""), str) else s, s.get("value"), s.get("ms_per_step"))
EOF
