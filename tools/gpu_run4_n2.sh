mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -3
for c in C3 C2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --no-e2e > gpurun_out/r2_n2b_$c.json 2> gpurun_out/r2_n2b_$c.err; echo "n2 $c rc=$?"
timeout 600 python bench.py --config $c --no-cpu-baseline > gpurun_out/r2c_$c.json 2> gpurun_out/r2c_$c.err
done
python - <<'PY'
import json
for f in ("r2_n2b_C3", "r2_n2b_C2", "r2c_C3", "r2c_C2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["step_ms"], "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"), (d.get("collective") or {}).get("coll_ms_per_batch"),
              {a: round(b["ms_per_step"], 3) for a, b in d["roofline"]["kernels"].items()} if "roofline" in d else "")
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
