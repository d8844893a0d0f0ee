# quick GPU check: full gpu test-suite + device-resident bench
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in ${GS_CONFIGS:-C2 C3}; do
timeout 600 python bench.py --config $c --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/b_$c.json 2> gpurun_out/b_$c.err
python - $c <<'PY'
import json, sys
c=sys.argv[1]
d=json.loads(open(f'gpurun_out/b_{c}.json').read().strip().splitlines()[-1])
print(c, d['value'], d['ms_per_step'], d['step_ms'], d['config'].get('instances_R'), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
PY
done
