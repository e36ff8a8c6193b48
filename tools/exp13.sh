python tools/three_check.py 24 16
python tools/three_check.py 25 8
python tools/three_check.py 26 3
python tools/three_check.py 27 1
python -m pytest tests/test_gpu_round2.py -q -m gpu -k "three_pass" 2>&1 | tail -3
