#!/bin/bash
mkdir -p gpurun_out
for p in fp32 fp16; do
python bench.py --workload formula --precision $p --steps 3 --warmup 3 > gpurun_out/r2b_bench_formula_$p.json 2> gpurun_out/r2b_bench_formula_$p.err
echo "formula $p exit $?"; tail -c 1800 gpurun_out/r2b_bench_formula_$p.json; tail -3 gpurun_out/r2b_bench_formula_$p.err
done
