# 2-GPU box: second-device-in-one-process test, N=2 bench (C3 default, C2), both with the warmed/timed collectives
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "second_device or autograd or prune" 2>&1 | tail -3
for c in C3 C2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c > gpurun_out/r2_n2_$c.json 2> gpurun_out/r2_n2_$c.err; echo "n2 $c rc=$?"
done
python - <<'PY'
import json
for f in ("r2_n2_C3", "r2_n2_C2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["step_ms"], "e2e", d.get("e2e"), d.get("collective"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
