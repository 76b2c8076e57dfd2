# Evidence pass on one B200: GPU tests, smoke, full bench line (with cpu / gpu baselines), reference arm, ncu launch list and
# ncu --set full capture of one training step (exports only: the .ncu-rep stays on the box).   usage: bash tools/gpu_evidence.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_$TAG.log
python bench.py > gpurun_out/r02_bench_n1_$TAG.json 2> gpurun_out/r02_bench_n1_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_$TAG.json 2> gpurun_out/r02_bench_reference_$TAG.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-profile --no-graph > gpurun_out/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:swin_|conv|proj_|fold_|lfq|wgrad|cls_|embed|rank1|anomaly|bce|adam' -o /tmp/step python tools/profile_step.py > gpurun_out/ncu_f_$TAG.log 2>&1; echo "set full rc=$?"
ncu -i /tmp/step.ncu-rep --page raw --csv > gpurun_out/r02_step_raw_$TAG.csv 2>/dev/null
ncu -i /tmp/step.ncu-rep --page source --print-source cuda,sass --csv -k 'regex:swin_attn_bwd_umma|conv96_wgrad|conv16_umma' > gpurun_out/r02_step_source_$TAG.csv 2>/dev/null
gzip -f gpurun_out/r02_launches_$TAG.csv gpurun_out/r02_step_raw_$TAG.csv gpurun_out/r02_step_source_$TAG.csv
ls -la gpurun_out/
