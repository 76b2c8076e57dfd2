mkdir -p gpurun_out
set -x
timeout -k 5 300 python -m pytest tests/test_bf16_gpu.py -x -q -k "tcgen05" > gpurun_out/r02_pytest_l.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_pytest_l.log
timeout -k 5 300 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_l.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['loss'])
for b in d['breakdown'][:16]: print("%-55s %.3f"%(b['op'],b['ms_per_step']))
P
