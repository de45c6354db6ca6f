timeout 900 python -m pytest tests/test_gpu_blocks.py -m gpu -q -p no:cacheprovider -k "ffn" 2>&1 | tail -3
for L in "" _nopf; do NBSS_LIB=nbss_b200/lib/libnbss_b200$L.so timeout 300 python tools/time_kernels.py 2>&1 | tail -1; done
NBSS_LIB=nbss_b200/lib/libnbss_b200_prof.so timeout 300 python tools/phase_profile.py --batch 32 2>&1 | sed -n "/ffn_bwd kernel 0/,/mhsa_fwd/p" | head -16
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_r2.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
