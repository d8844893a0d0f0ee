# Round-end check on the GPU box: what the driver runs (gpu test-suite, smoke, default bench for both arms) + the sanitizers.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
a = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
b = json.loads(open("gpurun_out/final_bench_ref.json").read().strip().splitlines()[-1])
print("ours", a["value"], a["e2e"]["value"], a["config"]["workload"], "| ref", b["value"], b["e2e"]["value"], "| ratio", round(a["value"] / b["value"], 2), round(a["e2e"]["value"] / b["e2e"]["value"], 2))
print("roofline", {k: a["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic")}, "cpu", a.get("cpu_baseline"), "launches", a.get("gpu_launches"), "clocks", a.get("clocks"))
PY
bash tools/gpu_sanitizer.sh
