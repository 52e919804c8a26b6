# one gpurun call: GPU test tier, bench lines (current build, previous build for A/B, rearrange), launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2b_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2b_gputests.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/r2b_gputests.log
tail -5 gpurun_out/r2b_gputests.log
timeout 300 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 600 gpurun_out/r2b_bench.json
if [ -f robogym_b200/librobogym_b200_prev.so ]; then RG_LIB=$PWD/robogym_b200/librobogym_b200_prev.so timeout 300 python bench.py > gpurun_out/r2b_bench_prevlib.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r2b_bench_prevlib.json'));print('prev lib', d['value'], d['ms_per_step'])"; fi
timeout 300 python bench.py --config rearrange_blocks > gpurun_out/r2b_bench_rearrange_blocks.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r2b_bench_rearrange_blocks.json'));print('rearrange', d['value'], d['ms_per_step'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_ncu_bench.log 2>&1; echo ncu rc=$?
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
