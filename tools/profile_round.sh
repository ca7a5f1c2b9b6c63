#!/bin/bash
# Collects the evidence of a round on the GPU box (run under gpurun, one GPU) into gpurun_out/:
#   bench_RR.json              the bench line (graph-replayed steps, CUDA events)
#   layer_times_RR.txt         per-call device times of an eager step, per (entry point, shape)
#   launches_RR.csv            ncu launch list of ONE eager training step (duration + DRAM bytes per launch;
#                              cold-cache, serialised) delimited by cudaProfilerStart/Stop (bench.py --ncu-window)
#   {igemm,dgrad,wgrad,image,bn,yolo}_full_RR.ncu-rep   --set full captures of a few launches of each hot kernel
R=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python tools/layer_times.py > gpurun_out/layer_times_$R.txt 2>&1
B="python bench.py --warmup 3 --ncu-window"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/launches_$R.csv $B > gpurun_out/launches_$R.log 2>&1
# forward igemm launches of the body (skip the small-channel stem), backward ones, wgrad
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_kernel \
    -s 20 -c 4 -o gpurun_out/igemm_full_$R -f $B > gpurun_out/igemm_full_$R.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_kernel \
    -s 130 -c 4 -o gpurun_out/dgrad_full_$R -f $B > gpurun_out/dgrad_full_$R.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_kernel \
    -s 30 -c 4 -o gpurun_out/wgrad_full_$R -f $B > gpurun_out/wgrad_full_$R.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_image \
    -c 2 -o gpurun_out/image_full_$R -f $B > gpurun_out/image_full_$R.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"bn_stats_apply|bn_bwd_stats_apply|bn_bwd_reduce" \
    -s 60 -c 6 -o gpurun_out/bn_full_$R -f $B > gpurun_out/bn_full_$R.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yolo_ -s 6 -c 6 \
    -o gpurun_out/yolo_loss_full_$R -f python tools/bench_yolo_loss.py > gpurun_out/yolo_loss_$R.log 2>&1
python tools/bench_yolo_loss.py > gpurun_out/yolo_loss_timing_$R.log 2>&1
bash tools/profile_extras.sh $R
