timeout 600 python -m pytest tests/test_disc.py tests/test_vgg.py -m gpu -q 2>&1 | tail -2
for v in 1 0; do echo "N128=$v"; B200SR_N128=$v timeout 300 python tools/bench_disc.py 2>&1 | tail -1 | cut -c1-160; B200SR_N128=$v timeout 300 python tools/bench_vgg.py 2>&1 | tail -1 | cut -c1-200; done
