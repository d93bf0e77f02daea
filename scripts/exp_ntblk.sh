# chunk-size experiment: template blocks of 128 per chunk on the overlap-save path (L2 residency of the A / P scratch)
for n in 1 2 4 8; do echo "NTBLK=$n"; FFTCONV_OS_NTBLK=$n python scripts/quick_time.py 1024 2>&1 | grep median; done
FFTCONV_OS_NTBLK=1 python scripts/path_time.py c2 3 1024 2>&1 | tail -8
