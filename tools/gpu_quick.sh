timeout 600 python tools/e2e_breakdown.py 2>&1 | tail -60
