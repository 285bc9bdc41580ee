#!/bin/bash
# End-of-round evidence in one gpurun call: full GPU suite, smoke, final bench line, ncu launch list of the same command, `ncu --set full` of the
# dominant kernel and of the kernels rewritten last. usage: bash tools/profile_final.sh r02
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -2 $out/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_final.json 2> $out/${tag}_bench_final.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $out/${tag}_launches_final.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $out/${tag}_launch_bench.log 2>&1; echo "launch list rc=$?"
for k in sbrt_inverse_multi fsdi_k4e_kernel lzp_walk_par_kernel lzp_spec_kernel lzi_parse_kernel; do
    timeout 400 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o $out/${tag}_full_$k \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $out/${tag}_full_$k.log 2>&1
    echo "full $k rc=$?"
done
