# developer A/B: landing-ring column pass in four-step (default on) and 2-D (developer knob) plans
for c in n22 n24 c5; do python tools/prof_case.py $c 10; done
echo "--- TFFT_NO_RING"
for c in n22 n24; do TFFT_DEVELOPER=1 TFFT_NO_RING=1 python tools/prof_case.py $c 10; done
echo "--- TFFT_RING_2D"
TFFT_DEVELOPER=1 TFFT_RING_2D=1 python tools/prof_case.py c5 10
