# developer A/B: MUFU inter-pass twiddles (default) against sincospif (-DTFFT_EXACT_TWIDDLE build)
for l in libtfft.so libtfft_exact.so; do echo "== $l"; for c in n16 n18 n20 n21 n22 n23 n24 c5; do TFFT_LIB=tensor-fft_b200/tfft/$l python tools/prof_case.py $c 10; done; done
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -m gpu -k "regression_guard or exec_twiddled or vs_fp64_oracle or three_pass_batched" 2>&1 | tail -3
