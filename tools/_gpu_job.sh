set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2g_gputests.log
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_raw.csv python tools/profile_step.py > gpurun_out/r2g_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_chain -s 4 -c 2 -o gpurun_out/r2g_chain -f python tools/profile_step.py > gpurun_out/r2g_ncu_c.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:wgrad3x3 -s 160 -c 2 -o gpurun_out/r2g_wgrad -f python tools/profile_step.py > gpurun_out/r2g_ncu_w.log 2>&1
tail -3 gpurun_out/r2g_gputests.log; cut -c1-300 gpurun_out/r2g_bench.json; tail -2 gpurun_out/r2g_bench.err
