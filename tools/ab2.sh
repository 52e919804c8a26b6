for cfg in "8 1" "4 2" "5 2"; do
  set -- $cfg
  RG_WARPS_PER_CTA=$1 RG_CTAS_PER_SM=$2 RG_LIB=$PWD/robogym_b200/librobogym_b200_ncon24.so timeout 200 python bench.py --steps 12 --warmup 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('warps/cta $1 ctas/sm $2', round(d['value']), round(d['ms_per_step'],2), d['config']['launch'])"
done
