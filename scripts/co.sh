for c in 100 75 50 35 25; do
CZB_HUF_CARVEOUT=$c python bench.py --frames 131072 --steps 3 --warmup 1 --distinct 512 --no-e2e --no-cpu-baseline > gpurun_out/co.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/co.json')); k=d['roofline']['kernel_ms_per_step']; print('carveout $c', 'huff=%.2f fse=%.2f exec=%.2f'%(k['huff'],k['fse'],k['exec']))"
done
