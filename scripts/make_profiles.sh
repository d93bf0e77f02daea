#!/bin/bash
# Run ON THE GPU BOX (gpurun -- bash scripts/make_profiles.sh rNN): ncu launch list of the bench command and
# one --set full capture of the hot kernels.  Outputs land in gpurun_out/; summarise them here with
# scripts/summarize_profiles.py (copies the judged summaries into profiles/).
R=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/${R}_ncu_bench.log 2>&1
# second call of the C2 workload (1000 templates = one chunk): os_data_fft, os_kern_fft, os_gemm, os_inverse
ncu --set full --clock-control none --import-source on -k regex:"os_data_fft|os_kern_fft|os_gemm|os_inverse" -s 4 -c 4 -f \
    -o gpurun_out/${R}_hot python scripts/ncu_os.py 3 1000 2 > gpurun_out/${R}_ncu_full.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_clocks.csv
tail -2 gpurun_out/${R}_ncu_full.log
