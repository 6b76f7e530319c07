python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"
