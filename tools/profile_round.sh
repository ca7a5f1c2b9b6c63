#!/bin/bash
# Collects the ncu evidence of a round on the GPU box (run under gpurun, one GPU):
#   gpurun_out/launches_rNN.csv   per-launch durations of one eager training step (ncu, cold-cache, serialised)
#   gpurun_out/igemm_full_rNN.ncu-rep / wgrad_full_rNN.ncu-rep   --set full captures of the conv kernels
#   gpurun_out/yolo_loss_full_rNN.ncu-rep   the fused YOLO loss kernel at 416^2 bs64 C=80
R=${1:-r01}
mkdir -p gpurun_out
export B200CV_CUDA_GRAPH=0
# one eager step = ~1300 launches; skip the first 4 steps (3 warm-up + 1) and list one step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5400 -c 1500 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline \
    > gpurun_out/launches_$R.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 300 -c 4 \
    -o gpurun_out/igemm_full_$R python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline \
    > gpurun_out/igemm_full_$R.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 100 -c 3 \
    -o gpurun_out/wgrad_full_$R python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline \
    > gpurun_out/wgrad_full_$R.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yolo_loss_nhwc -s 2 -c 2 \
    -o gpurun_out/yolo_loss_full_$R python tools/bench_yolo_loss.py > gpurun_out/yolo_loss_$R.log 2>&1
python tools/bench_yolo_loss.py > gpurun_out/yolo_loss_timing_$R.log 2>&1
ls -la gpurun_out
