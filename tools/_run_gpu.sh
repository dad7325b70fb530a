for sw in "10 3" "20 5" "40 5"; do set -- $sw
python bench.py --steps $1 --warmup $2 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('steps $1', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done
python tools/probe_host_e2e.py 2>&1 | tail -15
