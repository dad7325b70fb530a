set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2f_gputests.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_raw.csv python tools/profile_step.py > gpurun_out/r2f_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_chain -s 4 -c 2 -o gpurun_out/r2f_chain -f python tools/profile_step.py > gpurun_out/r2f_ncu_c.log 2>&1
tail -3 gpurun_out/r2f_gputests.log; cut -c1-600 gpurun_out/r2f_bench.json; tail -2 gpurun_out/r2f_bench.err
