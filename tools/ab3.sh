for w in 8 10; do
  RG_WARPS_PER_CTA=$w timeout 200 python bench.py --steps 12 --warmup 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('warps', '$w', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d['config']['launch'])"
done
