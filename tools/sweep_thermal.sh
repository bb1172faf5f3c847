for nt in 128 160 192 224 256; do
  echo "thermal NT=$nt: $(env SWALBE_NT=$nt python bench.py --workload thermal --steps 60 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['moments_only_mlups'])")"
done
