for l in libtfft_twtable.so libtfft.so libtfft_twtable.so libtfft.so; do
  echo "== $l"
  for c in n8 n9 n10 n12 n16 n18 n20; do env TFFT_LIB=$PWD/tensor-fft_b200/tfft/$l timeout 120 python tools/prof_case.py $c 20; done
done
