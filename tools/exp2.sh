# developer: cooperative prefetch / cluster A/B on the column-pass sizes
for v in "" "TFFT_NO_COOP=1" "TFFT_NO_CLUSTER=1" "TFFT_NO_CLUSTER=1 TFFT_NO_COOP=1"; do
  echo "== $v"
  for c in n16 n20 n21 n22 n23 n24 c5 n26; do env TFFT_DEVELOPER=1 $v timeout 120 python tools/prof_case.py $c 10; done
done
