# developer A/B: 64-column SWIZZLE_128B column tiles (default) against 16-column tiles (TFFT_NO_COL64)
for c in n16 n24; do python tools/prof_case.py $c 10; done
echo "--- TFFT_NO_COL64"
for c in n16 n24; do TFFT_DEVELOPER=1 TFFT_NO_COL64=1 python tools/prof_case.py $c 10; done
python tools/three_check.py 24 5
python tools/three_check.py 25 2
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -m gpu -k "regression_guard or vs_fp64_oracle or three_pass or four_step or tuner_knobs" 2>&1 | tail -3
