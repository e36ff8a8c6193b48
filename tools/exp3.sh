# developer A/B: current library vs libtfft_prev.so on the single-slot kernels
for l in libtfft_prev.so libtfft.so; do
  echo "== $l (no cluster)"
  for c in n8 n10 n12 n13 n15 n16 n20 n22 n24 c5; do env TFFT_LIB=$PWD/tensor-fft_b200/tfft/$l TFFT_DEVELOPER=1 TFFT_NO_CLUSTER=1 timeout 120 python tools/prof_case.py $c 10; done
done
echo "== libtfft.so cluster"
for c in n16 n22 n24 c5; do env TFFT_DEVELOPER=1 timeout 120 python tools/prof_case.py $c 10; done
