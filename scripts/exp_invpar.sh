# os_inverse_z: parity in the low bit of the warp index (one parity per scheduler partition), FFTCONV_OS_DBG=256
for d in 0 256 0 256; do FFTCONV_OS_DBG=$d python scripts/oneshot_time.py 1000 40 2>&1 | tail -1; done
for d in 0 256; do FFTCONV_OS_DBG=$d python scripts/config_time.py c4 2>&1 | grep -A5 "^\[c4\]"; done
