# round-end check on the GPU box: gpu test-suite, smoke(), default bench for both arms
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ['bench_default','bench_default_ref']:
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['step_ms'], 'e2e', d['e2e']['value'], d.get('gpu_launches'), d['clocks'], d.get('cpu_baseline',{}).get('value'))
PY
