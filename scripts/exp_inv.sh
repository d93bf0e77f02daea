echo Z-persist; python scripts/path_time.py c2 3 1000 2>&1 | grep "os_inverse\|median"
echo Z-oneshot; FFTCONV_OS_INV_PERSIST=0 python scripts/path_time.py c2 3 1000 2>&1 | grep "os_inverse\|median"
echo one-shot6; FFTCONV_OS_INV_Z=0 python scripts/path_time.py c2 3 1000 2>&1 | grep "os_inverse\|median"
ncu --set full --clock-control none --import-source on -k regex:"os_inverse" -s 2 -c 1 -f -o gpurun_out/r02d_inv python scripts/ncu_os.py 3 1000 3 > gpurun_out/r02d_ncu.log 2>&1
FFTCONV_OS_INV_PERSIST=0 ncu --set full --clock-control none --import-source on -k regex:"os_inverse" -s 2 -c 1 -f -o gpurun_out/r02e_inv python scripts/ncu_os.py 3 1000 3 > gpurun_out/r02e_ncu.log 2>&1
