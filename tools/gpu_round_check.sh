# One gpurun call that produces the round's evidence: GPU tests, smoke, the default bench, the ncu launch list of a short
# bench run and ncu --set full captures of the dominant kernels.  Usage: gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh TAG'
TAG=${1:-r1}
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests_$TAG.log; cat gpurun_out/gpu_tests_$TAG.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -6 | tee gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.err; head -c 300 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; head -c 400 gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --ba-iters 6 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1

timeout 300 ncu --set full --clock-control none --import-source on -k regex:fuse_joints -s 2 -c 1 -o gpurun_out/prof_fuse_joints_$TAG -f python tools/fusion_bench.py 200000 > gpurun_out/ncu_fuse_joints_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fuse_moments -s 2 -c 1 -o gpurun_out/prof_fuse_moments_$TAG -f python tools/fusion_bench.py 200000 > gpurun_out/ncu_fuse_moments_$TAG.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tri_|ba_|fuse_|ema_kernel|project|loss_|stats|post_|sg_|adam|so3|flag_counts|bone|pose_temporal|camera_|baseline" -c 2500 --csv --log-file gpurun_out/launches_ska_$TAG.csv python bench.py --steps 3 --warmup 3 --ba-iters 6 --no-cpu-baseline > gpurun_out/ncu_launch_ska_$TAG.log 2>&1

ls -la gpurun_out/*$TAG*
