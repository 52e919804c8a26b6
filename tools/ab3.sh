# A/B helper: warps per CTA sweep for a library variant ($1 = suffix, rest = warp counts)
v=$1; shift
for w in "$@"; do
  RG_LIB=$PWD/robogym_b200/librobogym_b200$v.so RG_WARPS_PER_CTA=$w timeout 200 python bench.py --steps 12 --warmup 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant [$v] warps', '$w', round(d['value']), round(d['ms_per_step'],2), d['config']['launch'], 'warn', d['config'].get('warn_bits'))"
done
