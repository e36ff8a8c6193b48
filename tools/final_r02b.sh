# round-2 (second session) validation run: GPU tests, bench line, C3 sweep, launch list, one capture of the ring kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t60_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/t60_pytest.log
python bench.py --steps 50 --warmup 5 > gpurun_out/t60_bench.log 2> gpurun_out/t60_bench.err; echo "bench rc $?" >> gpurun_out/t60_bench.err
python tools/sweep.py 28 > gpurun_out/t60_sweep.log 2>&1
cp gpurun_out/sweep.json gpurun_out/t60_sweep.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t60_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/t60_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ring -c 1 -f -o gpurun_out/t60_prof_ring python tools/prof_case.py n22 2 > gpurun_out/t60_ncu_ring.log 2>&1
