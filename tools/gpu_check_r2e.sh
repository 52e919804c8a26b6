# fourth gpurun call of the round: barrier-group A/B on the headline workload, smoke with 2 groups, and the `ncu --set full`
# capture of the first TIMED launch of the bench (launch 24 of `bench.py --steps 2 --warmup 3`: 20 settle + 3 warm-up before it)
mkdir -p gpurun_out
export RG_CPU_BASELINE_SECONDS=1
for g in 1 2 3 4 6; do
  RG_BAR_GROUPS=$g timeout 200 python bench.py --steps 40 --warmup 10 > gpurun_out/r2e_bench_groups$g.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r2e_bench_groups$g.json'));print('groups $g', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'warn', d['config']['warn_bits'])"
done
RG_BAR_GROUPS=2 timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rg_step_kernel -s 23 -c 1 -f -o gpurun_out/prof_r2c python bench.py --steps 2 --warmup 3 > gpurun_out/r2e_ncu.log 2>&1; echo ncu rc=$?
