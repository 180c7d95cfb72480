# One gpurun call that produces the round's evidence: GPU tests, smoke, the default bench, the ncu launch list of a short
# bench run and ncu --set full captures of the dominant kernels.  Usage: gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh TAG'
TAG=${1:-r1}
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests_$TAG.log; cat gpurun_out/gpu_tests_$TAG.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -6 | tee gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.err; head -c 300 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; head -c 400 gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --ba-iters 6 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
