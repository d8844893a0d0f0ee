mkdir -p gpurun_out
run() { # N config extra
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --config $2 $3 > gpurun_out/r2_n$1_$2.json 2> gpurun_out/r2_n$1_$2.err; echo "n$1 $2 rc=$?"
}
run 8 C3 ""
run 8 C2 "--no-e2e"
run 4 C3 "--no-e2e"
python - <<'PY'
import json
for f in ("r2_n8_C3", "r2_n8_C2", "r2_n4_C3"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["step_ms"], "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"), (d.get("collective") or {}).get("coll_ms_per_batch"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-3000:])
PY
