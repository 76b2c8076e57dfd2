mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_h.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_gpu_h.log
python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_h.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'])
for b in d['breakdown'][:40]: print("%-55s %.3f"%(b['op'],b['ms_per_step']))
P
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:swin_fwd_umma|swin_attn_bwd_umma|swin_mlp_bwd_umma|wgrad_tc_kernel|proj_bwd_scalar|conv_tc16_kernel|conv_tc_kernel|fold_pad|lfq_|conv16_umma|conv96_umma' -o gpurun_out/r02_step_h python tools/profile_step.py > gpurun_out/ncu_step_h.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_step_h.log
ls -la gpurun_out/r02_step_h.ncu-rep
