set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/r02b_pytest.log
timeout 600 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
cat gpurun_out/r02b_bench.json | head -c 6000
bash scripts/make_profiles.sh r02b
