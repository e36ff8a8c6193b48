# developer A/B: bench the default library against alternative builds given as arguments
for l in libtfft.so "$@"; do
TFFT_LIB=tensor-fft_b200/tfft/$l timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$l', d['ms_per_step'], d['roofline']['frac'])"
done
