# usage: bash tools/sweep.sh "128 192 256" "64 128 256"
for nt in ${1:-128 192 256}; do for rmax in ${2:-128 256}; do
  echo "NT=$nt RMAX=$rmax: $(SWALBE_NT=$nt SWALBE_RMAX=$rmax python bench.py --steps ${STEPS:-150} --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['moments_only_mlups'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))")"
done; done
