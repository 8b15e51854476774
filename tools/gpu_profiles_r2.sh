#!/bin/bash
# round-2 evidence: ncu --set full of the dominant kernels (one launch each) + the launch list of bench.py.
# Each ncu line follows a plain run of the same command; run the blocks in separate gpurun calls when GPU time is short.
set -x
mkdir -p gpurun_out
export KC_TIME_T=4
python tools/time_knode.py 18944 > gpurun_out/plain_knode.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:kc_knode_tc_fwd_kernel -s 1 -c 1 -o gpurun_out/r02_prof_knode_tc_fwd python tools/time_knode.py 18944 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kc_knode_tc_bwd_kernel -s 1 -c 1 -o gpurun_out/r02_prof_knode_tc_bwd python tools/time_knode.py 18944 > gpurun_out/ncu2.log 2>&1
}
unset KC_TIME_T
python tools/prof_train.py 1024 3 > gpurun_out/plain_train.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kc_train_tc3_kernel -s 3 -c 1 -o gpurun_out/r02_prof_train_tc3 python tools/prof_train.py 1024 3 > gpurun_out/ncu3.log 2>&1
python tools/prof_rollout.py 4096 100 f32 2 > gpurun_out/plain_rollout.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kc_rollout_wide_lin_kernel -s 1 -c 1 -o gpurun_out/r02_prof_rollout_lin python tools/prof_rollout.py 4096 100 f32 2 > gpurun_out/ncu4.log 2>&1
python tools/prof_estimate.py > gpurun_out/plain_est.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kc_estimate_kernel -s 2 -c 1 -o gpurun_out/r02_prof_estimate python tools/prof_estimate.py > gpurun_out/ncu5.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kc_ -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu6.log 2>&1
tail -n 2 gpurun_out/plain_*.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_bench.csv
