mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/t10_bench.log 2> gpurun_out/t10_bench.err
echo "--- 2^24 default vs three-pass" 
python tools/prof_case.py n24 10
TFFT_DEVELOPER=1 TFFT_THREEPASS_LG=24 python tools/prof_case.py n24 10
echo "--- C5 ybits default vs 2"
python tools/prof_case.py c5 10
TFFT_DEVELOPER=1 TFFT_2D_YBITS=2 python tools/prof_case.py c5 10
echo "--- 2^22 / 2^23 lg1 alternatives"
for lg1 in 10 11 12; do TFFT_DEVELOPER=1 TFFT_FOURSTEP_LG1=$lg1 python tools/prof_case.py n22 10; done
for lg1 in 11 12; do TFFT_DEVELOPER=1 TFFT_FOURSTEP_LG1=$lg1 python tools/prof_case.py n23 10; done
