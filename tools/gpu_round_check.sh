set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests_r1z.log; cat gpurun_out/gpu_tests_r1z.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -6 | tee gpurun_out/smoke_r1z.log
timeout 600 python bench.py > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; tail -c 600 gpurun_out/bench_r1z.err; head -c 300 gpurun_out/bench_r1z.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ba_calib_linearize -s 2 -c 1 -o gpurun_out/prof_ba_calib_r1z -f python tools/ba_bench.py c3 100000 3 --calib > gpurun_out/ncu_ba_calib_r1z.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fuse_frames -s 2 -c 1 -o gpurun_out/prof_fuse_r1z -f python tools/fusion_bench.py 200000 > gpurun_out/ncu_fuse_r1z.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ema_kernel -s 2 -c 1 -o gpurun_out/prof_ema_r1z -f python tools/fusion_bench.py 1000000 > gpurun_out/ncu_ema_r1z.log 2>&1
ls -la gpurun_out/*r1z*
