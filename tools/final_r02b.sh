# round-2 (second session) validation run: GPU tests, bench line, C3 sweep, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t70_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/t70_pytest.log
python bench.py --steps 50 --warmup 5 > gpurun_out/t70_bench.log 2> gpurun_out/t70_bench.err; echo "bench rc $?" >> gpurun_out/t70_bench.err
python tools/sweep.py 28 > gpurun_out/t70_sweep.log 2>&1
cp gpurun_out/sweep.json gpurun_out/t70_sweep.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t70_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/t70_ncu_list.log 2>&1
