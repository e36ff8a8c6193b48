# developer A/B: 32-column SWIZZLE_64B column tiles (default for 512-point column passes) against 16-column tiles
for c in n17 n18 n19; do python tools/prof_case.py $c 10; done
echo "--- TFFT_NO_COL64"
for c in n17 n18 n19; do TFFT_DEVELOPER=1 TFFT_NO_COL64=1 python tools/prof_case.py $c 10; done
python tools/three_check.py 25 4
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -m gpu -k "regression_guard or vs_fp64_oracle or three_pass or four_step or tuner_knobs or real_reference" 2>&1 | tail -3
