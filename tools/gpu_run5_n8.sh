# 8-GPU box: what the driver's scaling run does (N = 1, 2, 4, 8 on the default config, both arms at N = 1) + C2 / C4 / C5 at 8 ranks
mkdir -p gpurun_out
run() { # N config extra tag
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --config $2 $3 > gpurun_out/r2s_n$1_$2.json 2> gpurun_out/r2s_n$1_$2.err; echo "n$1 $2 rc=$?"
}
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s_n1_C3.json 2> gpurun_out/r2s_n1_C3.err
run 2 C3 "--no-cpu-baseline"
run 4 C3 "--no-cpu-baseline"
run 8 C3 "--no-cpu-baseline"
run 8 C2 "--no-e2e"
run 8 C4 "--no-e2e"
run 8 C5 "--no-e2e"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --impl reference > gpurun_out/r2s_n8_ref.json 2> gpurun_out/r2s_n8_ref.err; echo "ref n8 rc=$?"
python - <<'PY'
import json
base = None
for f in ("r2s_n1_C3", "r2s_n2_C3", "r2s_n4_C3", "r2s_n8_C3", "r2s_n8_C2", "r2s_n8_C4", "r2s_n8_C5", "r2s_n8_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        if f == "r2s_n1_C3": base = d["value"]
        eff = round(d["value"] / (d["n_gpus"] * base), 3) if base and f.endswith("C3") else None
        c = d.get("collective") or {}
        print(f, d["value"], d["ms_per_step"], d["step_ms"]["median"], "eff", eff, "e2e", (d.get("e2e") or {}).get("value"), c.get("coll_ms_per_batch"), c.get("busbw_GBps"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
