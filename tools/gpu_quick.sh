for c in C4 C5; do
timeout 600 python bench.py --config $c --no-cpu-baseline --steps 8 > gpurun_out/bench_ours_$c.json 2> gpurun_out/bench_ours_$c.err; echo "ours $c rc=$?"
timeout 600 python bench.py --config $c --impl reference --steps 8 > gpurun_out/bench_ref_$c.json 2> gpurun_out/bench_ref_$c.err; echo "ref $c rc=$?"
done
python - <<'PY'
import json
for f in ['ours_C4','ref_C4','ours_C5','ref_C5']:
    try:
        d=json.loads(open(f'gpurun_out/bench_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['value'], d['config']['workload'], d['config'].get('instances_R'))
        if 'roofline' in d: print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e); print(open(f'gpurun_out/bench_{f}.err').read()[-1500:])
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
