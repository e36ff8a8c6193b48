# developer: tfft_exec_host chunk size sweep (TFFT_HOST_CHUNK_MB is a developer knob)
for mb in 4 8 16 32; do
TFFT_DEVELOPER=1 TFFT_HOST_CHUNK_MB=$mb timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk_mb', $mb, 'e2e_ms', d['e2e']['ms_per_step'], 'ceiling_ms', d['e2e'].get('pcie_ceiling_ms'))"
done
