# round-2 GPU check #1: full gpu test-suite, default bench (C3) both arms, C2, binning-plan experiment, sanitizers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py > gpurun_out/r2_bench_C3.json 2> gpurun_out/r2_bench_C3.err; echo "bench C3 rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_ref_C3.json 2> gpurun_out/r2_bench_ref_C3.err; echo "bench ref C3 rc=$?"
timeout 600 python bench.py --config C2 --no-cpu-baseline > gpurun_out/r2_bench_C2.json 2> gpurun_out/r2_bench_C2.err; echo "bench C2 rc=$?"
for v in 1 2; do
GSB_BIN_PER_SM=$v timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_C3_persm$v.json 2> gpurun_out/r2_bench_C3_persm$v.err
done
python - <<'PY'
import json
for f in ("r2_bench_C3", "r2_bench_ref_C3", "r2_bench_C2", "r2_bench_C3_persm1", "r2_bench_C3_persm2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        k = {a: round(b["ms_per_step"], 3) for a, b in d.get("roofline", {}).get("kernels", {}).items()}
        print(f, d["value"], d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"), k)
    except Exception as e:
        print(f, "ERR", e)
PY
bash tools/gpu_sanitizer.sh
