python tools/three_check.py 24 16
TFFT_DEVELOPER=1 TFFT_THREEPASS_LOOP=1 python tools/three_check.py 24 16
TFFT_DEVELOPER=1 TFFT_THREEPASS_LG=25 python tools/three_check.py 24 16
python tools/three_check.py 24 1
python tools/three_check.py 25 8
TFFT_DEVELOPER=1 TFFT_THREEPASS_LOOP=1 python tools/three_check.py 25 8
python tools/three_check.py 26 3
