CHAIN_SWEEP="3" timeout 600 python tools/chain_check.py 512 2>&1 | grep -E "vs unchained|FAILED|median"
python bench.py --batch 512 --no-cpu-baseline > gpurun_out/r2_x_bench_b512.json 2> gpurun_out/r2_x_bench_b512.err; tail -c 300 gpurun_out/r2_x_bench_b512.err
