if [ -z "$NOTEST" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -m gpu -x -q -k "plane or c3" 2>&1 | tail -3; fi
for v in ${VARS:-0 1 2 3 4 5}; do echo "BP_CT=$v"; FFTCONV_BP_CT=$v timeout 300 python scripts/c3_time.py 16 2>&1 | tail -5 | grep -v repad; done
