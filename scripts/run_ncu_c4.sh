# --set full capture of the config-4 kernels (one group of 4 images = 1156 tiles, the unit fftconv_conv_batch repeats 16 times)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"os_data_fft|os_gemm|os_inverse" -s 3 -c 3 -f \
    -o gpurun_out/r02e_c4 python scripts/ncu_c4.py 4 > gpurun_out/r02e_c4_ncu.log 2>&1
tail -3 gpurun_out/r02e_c4_ncu.log
