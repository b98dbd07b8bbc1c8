for b in 0 1; do
echo "DEVO_CORR_BWD_BULK=$b"
DEVO_CORR_BWD_BULK=$b timeout 300 python -m pytest tests/test_gpu_corr.py tests/test_parity_vs_reference_ext.py -x -q -m gpu -p no:cacheprovider -k "backward" 2>&1 | tail -1
DEVO_CORR_BWD_BULK=$b timeout 200 python tools/corr_bwd_timing.py 2>&1 | tail -2
done
