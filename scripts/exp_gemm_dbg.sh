for d in 0 1 4 8 5 13; do echo "OS_DBG=$d"; FFTCONV_OS_DBG=$d timeout 200 python scripts/c5_time.py 2048 2>&1 | grep -A8 "one_call=True" | grep "os_gemm\|os_inverse\|one_call"; done
