timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python tools/bench_tools.py > gpurun_out/bench_tools.jsonl 2> gpurun_out/bench_tools.err; cat gpurun_out/bench_tools.jsonl; tail -3 gpurun_out/bench_tools.err
