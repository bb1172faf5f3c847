# theta-field flavours: geometry chosen and rate, TF (strict, compiled for a field) vs OPTS (run-time options)
run() { echo "$1: $(env $1 SWALBE_DEBUG=1 python bench.py --workload thermal_moving --steps 60 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep -E 'geometry.*thermal=1|"metric"' | sed -E 's/.*(NT=[0-9]+ W=[0-9]+ strips=[0-9]+ rows.CTA=[0-9]+ chunks=[0-9]+ CTAs.SM=[0-9]+).*/\1/; s/.*"value": ([0-9.]+).*/\1 MLUPS/' | sort -u | tr '\n' ' ')"; }
run SWALBE_TF=1
run SWALBE_TF=0
run "SWALBE_TF=1 SWALBE_NT=128"
run "SWALBE_TF=0 SWALBE_NT=128"
run "SWALBE_TF=1 SWALBE_NT=160"
