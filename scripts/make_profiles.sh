#!/bin/bash
# Run ON THE GPU BOX (gpurun -- bash scripts/make_profiles.sh rNN): ncu launch list of the bench command and
# one --set full capture of the hot kernels.  Outputs land in gpurun_out/; summarise them here with
# scripts/summarize_profiles.py (copies the judged summaries into profiles/).
R=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${R}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tile16_conv|tile16_c2r|tile16_kern_hpass" -s 9 -c 3 -f \
    -o gpurun_out/${R}_hot python scripts/quick_time.py 1000 0 > gpurun_out/${R}_ncu_full.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_clocks.csv
tail -2 gpurun_out/${R}_ncu_full.log
