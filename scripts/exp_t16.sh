timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py > gpurun_out/r02g_racecheck.txt 2>&1; grep -c "Race reported" gpurun_out/r02g_racecheck.txt; grep "Race reported\|RACECHECK SUMMARY" gpurun_out/r02g_racecheck.txt | cut -c1-220 | head
python bench.py --workload c1 --no-cpu --no-configs --no-extras 2>/dev/null | tail -1 | cut -c1-300
