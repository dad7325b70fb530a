for i in 1 2; do for n in 1 2 3 4; do
B200SR_WGRAD_STREAMS=$n python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('streams $n', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done; done
python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "gradients or golden" 2>&1 | tail -2
