#!/bin/bash
# Collects the ncu evidence of a round on the GPU box (run under gpurun, one GPU):
#   gpurun_out/launches_rNN.csv   per-launch durations of ONE eager training step (ncu, cold-cache, serialised);
#                                 the step is delimited by cudaProfilerStart/Stop (bench.py --ncu-window)
#   gpurun_out/{igemm,wgrad}_full_rNN.ncu-rep   --set full captures of a few conv launches of that step
#   gpurun_out/yolo_loss_full_rNN.ncu-rep       the fused YOLO loss kernels at 416^2 bs64 C=80
R=${1:-r01}
mkdir -p gpurun_out
B="python bench.py --warmup 3 --ncu-window"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$R.csv $B > gpurun_out/launches_$R.log 2>&1
# forward igemm launches of the body (skip the first 20 = small-channel stem), a few backward ones, a few wgrad
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_kernel \
    -s 20 -c 6 -o gpurun_out/igemm_full_$R -f $B > gpurun_out/igemm_full_$R.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_kernel \
    -s 130 -c 6 -o gpurun_out/dgrad_full_$R -f $B > gpurun_out/dgrad_full_$R.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_kernel \
    -s 30 -c 6 -o gpurun_out/wgrad_full_$R -f $B > gpurun_out/wgrad_full_$R.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yolo_ -s 6 -c 6 \
    -o gpurun_out/yolo_loss_full_$R -f python tools/bench_yolo_loss.py > gpurun_out/yolo_loss_$R.log 2>&1
python tools/bench_yolo_loss.py > gpurun_out/yolo_loss_timing_$R.log 2>&1
ls -la gpurun_out
