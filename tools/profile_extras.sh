#!/bin/bash
# Evidence for the SURVEY 8f rows (inference joint = BASELINE config 5, optimizer step) and the reproducibility probe.
# Called by profile_round.sh; can run alone.  Ends by condensing every .ncu-rep ON THE BOX into gpurun_out/profiles_RR/
# (tools/make_profile_summary.py) and pruning the raw reports so that gpurun_out/ stays under gpurun's 64 MiB limit.
R=${1:-r01}
mkdir -p gpurun_out
# inference joint (BASELINE config 5), optimizer step, reproducibility probe
python tools/bench_pipeline.py --out gpurun_out/pipeline_$R.json > /dev/null 2> gpurun_out/pipeline_$R.err
python tools/bench_optim.py > gpurun_out/optim_$R.json 2> gpurun_out/optim_$R.err
python tools/determinism_probe.py 2>/dev/null | grep "rel spread" > gpurun_out/determinism_$R.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"adam_multi|sgd_multi" -s 3 -c 2 \
    -o gpurun_out/optim_full_$R -f python tools/bench_optim.py --iters 2 > gpurun_out/optim_full_$R.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"detect_nms|crop_resize|detect_compact" -s 3 -c 3 \
    -o gpurun_out/detect_full_$R -f python tools/bench_pipeline.py --iters 2 --warmup 1 > gpurun_out/detect_full_$R.log 2>&1
python tools/make_profile_summary.py $R gpurun_out/profiles_$R > gpurun_out/make_profile_summary_$R.log 2>&1
# keep the raw reports only while the directory fits (largest first out)
while [ "$(du -sm gpurun_out | cut -f1)" -ge 60 ]; do
    big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; rm -f "$big"
done
ls -la gpurun_out
