for L in 100 256 1024 2048 4096 8192 16384; do
  steps=200; [ $L -le 1024 ] && steps=2000
  echo "L=$L: $(python bench.py --L $L --steps $steps --warmup 10 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MLUPS',d['value'],'ms/step',d['ms_per_step'],'lazy',d.get('moments_only_mlups'),'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])")"
done
echo "thermal 8192: $(python bench.py --workload thermal --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MLUPS',d['value'],'ms/step',d['ms_per_step'],'lazy',d.get('moments_only_mlups'),'e2e',d['e2e']['value'])")"
