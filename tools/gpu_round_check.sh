# One gpurun call that produces the round's evidence: GPU tests, smoke, the default bench, the reference arm, the ncu launch
# list of a short bench run and ncu --set full captures of the dominant kernels.
# Usage: gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh TAG'
TAG=${1:-r02}
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests_$TAG.log; cat gpurun_out/gpu_tests_$TAG.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -6 | tee gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.err; head -c 300 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; head -c 400 gpurun_out/bench_ref_$TAG.json
timeout 300 python tools/tri_bench.py 2>&1 | tee gpurun_out/tri_bench_$TAG.log
timeout 300 python tools/ba_reg_bench.py 2>&1 | tee gpurun_out/ba_reg_bench_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tri_|ba_|reg_|fuse_|rigid_|ema_kernel|project|loss_|stats|post_|sg_|adam|so3|flag_counts|bone|pose_temporal|camera_|baseline" -c 3000 --csv --log-file gpurun_out/launches_ska_$TAG.csv python bench.py --steps 3 --warmup 3 --ba-iters 6 --no-cpu-baseline > gpurun_out/ncu_launch_ska_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_ska_$TAG.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
full() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c 1 -o gpurun_out/prof_${name}_$TAG -f "$@" > gpurun_out/ncu_${name}_$TAG.log 2>&1
  # gpurun brings back at most 64 MiB: summarise on the box (raw + source pages), keep the raw report of the headline kernels only
  timeout 120 python tools/ncu_summary.py gpurun_out/prof_${name}_$TAG.ncu-rep > gpurun_out/${TAG}_${name}_ncu_full_summary.txt 2>&1
  case $name in tri_cta_config2|tri_cta_vp_8view) ;; *) rm -f gpurun_out/prof_${name}_$TAG.ncu-rep ;; esac
}
full tri_cta_config2 "tri_kernel_cta" 5 python tools/tri_bench.py c2
full tri_cta_vp_8view "tri_kernel_cta_vp" 5 python tools/tri_bench.py v8
full tri_cta_vp_config4 "tri_kernel_cta_vp" 5 python tools/tri_bench.py c4
full tri_cta_frames "tri_kernel_cta_frames" 5 python tools/tri_bench.py frames
full ba_wide_config5 "ba_linearize_wide_kernel" 2 python tools/ba_bench.py c5s 125000 3
full ba_tc_config5 "ba_linearize_tc_kernel" 2 python tools/ba_bench.py c5s 125000 3 --tc
full reg_matvec "reg_matvec_kernel" 2 python tools/ba_reg_bench.py 100000 17 2b full
full reg_precond "reg_precond_kernel" 2 python tools/ba_reg_bench.py 100000 17 2b full
full reg_linearize "reg_linearize_kernel" 1 python tools/ba_reg_bench.py 100000 17 2b full
full reg_pt_matvec "reg_pt_matvec_kernel" 2 python tools/ba_reg_bench.py 100000 17 2b pose_only
ls -la gpurun_out/*$TAG*
