for i in 1 2; do
B200SR_LIB=$GRAFT_REPO_ROOT/sr_gan_fd_b200/libb200sr_base.so python tools/quick_bench.py 2>&1 | grep "step" | sed 's/^/base    /'
python tools/quick_bench.py 2>&1 | grep "step" | sed 's/^/tailall /'
done
python tools/bench_disc.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('disc default', d.get('b200') or d)" 2>&1 | cut -c1-300
B200SR_WGRAD_STREAMS=4 B200SR_WGRAD_SMS=99 python tools/bench_disc.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('disc 4x99   ', d.get('b200') or d)" 2>&1 | cut -c1-300
B200SR_WGRAD_STREAMS=4 python tools/bench_disc.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('disc 4xall  ', d.get('b200') or d)" 2>&1 | cut -c1-300
