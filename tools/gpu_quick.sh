timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2 3 4 5 6; do
GS_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_C2_$i.json 2> gpurun_out/b_C2_$i.err
grep -E "^step_ms|^alloc_retries" gpurun_out/b_C2_$i.err
done
for i in 1; do
GS_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline --config C3 > gpurun_out/b_C3_$i.json 2> gpurun_out/b_C3_$i.err
grep -E "^step_ms|^alloc_retries" gpurun_out/b_C3_$i.err
done
python - <<'PY'
import json
for c in ['C2_1','C2_2','C2_3','C2_4','C2_5','C2_6','C3_1']:
    d=json.loads(open(f'gpurun_out/b_{c}.json').read().strip().splitlines()[-1])
    print(c, d['value'], d['ms_per_step'], d['step_ms'], d['e2e']['value'])
PY
