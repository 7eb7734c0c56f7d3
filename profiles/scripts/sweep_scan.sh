for cfg in "0 8" "0.3 6" "0.4 6" "0.4 8" "0.5 8" "0.5 10" "0.6 8" "0.6 10" "1.0 8"; do
  set -- $cfg
  MB200_SCAN_GATHER_FRAC=$1 MB200_SCAN_GATHER_WARPS=$2 python bench.py --workload scan --nseq 1000000 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('frac $1 warps $2:', round(d['value']/1e6,1), 'Mbp/s', round(d['ms_per_step'],1), 'ms', d['checks']['counts_sum'][0])"
done
