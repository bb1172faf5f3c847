# A/B two builds of the library on the same box: tools/ab.sh <alt.so> [bench args...]
alt=$1; shift
for i in 1 2 3; do
  echo "main: $(python bench.py --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'])")"
  echo "alt : $(SWALBE_B200_SO=$alt python bench.py --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'])")"
done
