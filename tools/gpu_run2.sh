mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for i in 1 2; do
GS_BENCH_DEBUG=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2b_C3_$i.json 2> gpurun_out/r2b_C3_$i.err; grep -E "^step_ms|alloc_retries" gpurun_out/r2b_C3_$i.err
done
timeout 600 python bench.py --config C2 --no-cpu-baseline > gpurun_out/r2b_C2.json 2> gpurun_out/r2b_C2.err
python - <<'PY'
import json
for f in ("r2b_C3_1", "r2b_C3_2", "r2b_C2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        k = {a: round(b["ms_per_step"], 3) for a, b in d.get("roofline", {}).get("kernels", {}).items()}
        print(f, d["value"], d["ms_per_step"], d["step_ms"], "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"), k)
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 python tools/e2e_breakdown.py 2>&1 | head -12
timeout 600 python tools/host_profile.py C2 2>&1 | head -60
