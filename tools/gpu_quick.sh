# quick GPU check: full gpu test-suite + bench at $GS_CONFIGS (default "C3 C2"); $GS_BENCH_ARGS are appended (e.g. --no-e2e)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
for c in ${GS_CONFIGS:-C3 C2}; do
timeout 600 python bench.py --config $c --no-cpu-baseline $GS_BENCH_ARGS > gpurun_out/q_$c.json 2> gpurun_out/q_$c.err
python - $c <<'PY'
import json, sys
c = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/q_{c}.json').read().strip().splitlines()[-1])
    print(c, d['value'], d['ms_per_step'], d['step_ms'], 'e2e', d.get('e2e', {}).get('value'), d.get('e2e', {}).get('ms_per_step'), {k: round(v['ms_per_step'], 3) for k, v in d['roofline']['kernels'].items()})
except Exception as e:
    print(c, 'ERR', e); print(open(f'gpurun_out/q_{c}.err').read()[-1500:])
PY
done
