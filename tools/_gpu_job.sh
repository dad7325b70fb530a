ls /root/reference 2>&1 | head -2
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
