timeout 900 python -m pytest tests/test_gpu_tools.py -x -q -k loss 2>&1 | tail -12
timeout 600 python tools/bench_loss.py 2>&1 | tail -3
