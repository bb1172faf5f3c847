# usage: bash tools/scaling.sh  (on an 8-GPU box)
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --warmup 10 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['scaling'], d['config']['grid'], 'MLUPS', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'] if d['e2e'] else None, 'frac', d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))"; }
echo "--- weak, film 8192^2 per GPU"; for n in 8 4 2; do run $n --steps 200; done
python bench.py --steps 200 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(1, 'weak', d['config']['grid'], 'MLUPS', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"
echo "--- weak, thermal 8192^2 per GPU"; run 8 --steps 100 --workload thermal
echo "--- strong, film 16384^2 total"; for n in 8 4 2; do run $n --steps 100 --L 16384 --scaling strong; done
python bench.py --steps 100 --L 16384 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(1, 'strong', d['config']['grid'], 'MLUPS', d['value'], 'ms/step', d['ms_per_step'])"
