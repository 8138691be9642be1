for mb in 512 1024 3072 8192; do
CZB_HOST_CHUNK_MB=$mb python bench.py --frames 65536 --steps 1 --warmup 1 --distinct 512 --e2e-frames 262144 --e2e-steps 2 --no-cpu-baseline > gpurun_out/e2e.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/e2e.json')); e=d['e2e']; print('chunk_mb $mb', 'e2e GB/s=%.1f ms=%.0f'%(e['value'],e['ms_per_step']))"
done
