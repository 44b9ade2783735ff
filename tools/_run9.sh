set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload timesformer_fwd --steps 10 --warmup 3 > gpurun_out/r02w_timesformer_fwd.json 2> gpurun_out/r02w_timesformer_fwd.err
python bench.py --steps 10 --warmup 3 --no-torch-baseline --no-cpu-baseline > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err
ALPRO_ATTN_PS=0 python bench.py --steps 10 --warmup 3 --no-torch-baseline --no-cpu-baseline > gpurun_out/r02w_bench_nops.json 2> gpurun_out/r02w_bench_nops.err
python - <<'PY'
import json
for f in ("gpurun_out/r02w_timesformer_fwd.json", "gpurun_out/r02w_bench.json", "gpurun_out/r02w_bench_nops.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("power"), d["clocks"], d["roofline"]["frac"], d.get("losses"))
PY
tail -3 gpurun_out/r02w_bench.err
