# A/B helper: bench.py for each library variant given as argument (path suffixes), two runs each
for v in "$@"; do
  for rep in 1 2; do
    RG_LIB=$PWD/robogym_b200/librobogym_b200$v.so timeout 200 python bench.py --steps 12 --warmup 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant', '[$v]', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d['config']['launch'])"
  done
done
