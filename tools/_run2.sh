set -x
python tools/energy_probe.py gpurun_out/r02q_energy_probe.md 2>&1 | tail -40
python bench.py --steps 10 --warmup 3 --no-torch-baseline --no-cpu-baseline > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['power'], d['clocks'])
"
