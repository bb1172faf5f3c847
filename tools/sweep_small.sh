# geometry sweep at the mid-size configs (C2 1024^2, 2048^2): chosen geometry vs forced NT x rows/CTA
for L in 1024 2048; do
  echo "L=$L default: $(env SWALBE_DEBUG=1 python bench.py --L $L --steps 1000 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | grep -m1 'geometry' | cut -c1-160)"
  echo "L=$L default: $(python bench.py --L $L --steps 1000 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])")"
  for nt in 128 160 192 224 256; do for rows in 16 24 32 48 64; do
    echo "L=$L NT=$nt ROWS=$rows: $(env SWALBE_NT=$nt SWALBE_ROWS=$rows python bench.py --L $L --steps 1000 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])")"
  done; done
done
