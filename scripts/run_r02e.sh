set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02e_pytest.log
timeout 600 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
cat gpurun_out/r02e_bench.json | head -c 3000
bash scripts/make_profiles.sh r02e
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
H=cuda-fft-convolution_b200/harness/fftconv_bench
( $H --config c1 --check 64; $H --config c2 --check 16; $H --config c2 --host; $H --config c3 --iters 3; $H --config c4 --iters 3; $H --config c5 --iters 2 ) > gpurun_out/r02e_harness.txt 2>&1; echo "harness rc=$?"
tail -20 gpurun_out/r02e_harness.txt
