# developer: effect of the 3R-column DFT matrix (two 16K-element CTAs per SM for 2048-point units)
echo "== default"; for c in c2 n11 n12 n22 n23 c5; do timeout 120 python tools/prof_case.py $c 10; done
echo "== 2048 as 16K units"; TFFT_DEVELOPER=1 TFFT_U2048_16K=1 timeout 120 python tools/prof_case.py n11 10
echo "== lg1=11"; for c in n22 n23; do TFFT_DEVELOPER=1 TFFT_FOURSTEP_LG1=11 timeout 120 python tools/prof_case.py $c 10; done
echo "== lg1=11, 8 columns for 2048"; for c in n22 n23; do TFFT_DEVELOPER=1 TFFT_FOURSTEP_LG1=11 TFFT_COL2048_U8=1 timeout 120 python tools/prof_case.py $c 10; done
echo "== c5 ybits 2"; TFFT_DEVELOPER=1 TFFT_2D_YBITS=2 timeout 120 python tools/prof_case.py c5 10
