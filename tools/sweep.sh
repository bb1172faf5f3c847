# usage: bash tools/sweep.sh "128 192 256" "64 128 256" [L] [steps]
L=${3:-8192}; STEPS=${4:-${STEPS:-150}}
for nt in ${1:-128 192 256}; do for rmax in ${2:-128 256}; do
  echo "L=$L NT=$nt RMAX/ROWS=$rmax: $(env SWALBE_NT=$nt SWALBE_RMAX=$rmax ${ROWS:+SWALBE_ROWS=$rmax} python bench.py --L $L --steps $STEPS --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['moments_only_mlups'])")"
done; done
