mkdir -p gpurun_out
TAG=${1:-x}
timeout -k 5 300 python -m pytest tests/test_bf16_gpu.py -x -q -k "tcgen05" > gpurun_out/r02_pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_pytest_$TAG.log
timeout -k 5 300 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_$TAG.json 2> gpurun_out/r02_bench_$TAG.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_$TAG.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'])
for b in d['breakdown'][:22]: print("%-55s %.3f"%(b['op'],b['ms_per_step']))
P
