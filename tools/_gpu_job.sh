timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-400
