# marching kernel vs tile kernel on small lattices (us per step, MLUPS)
for L in 100 256 512 768 1024 2048; do
  steps=2000; [ $L -ge 1024 ] && steps=500
  for tm in 0 100000000; do
    echo "L=$L TILE_MAX=$tm: $(env SWALBE_TILE_MAX=$tm python bench.py --L $L --steps $steps --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], 'MLUPS', round(d['ms_per_step']*1e3,2), 'us/step  moments-only', d.get('moments_only_mlups'))")"
  done
done
