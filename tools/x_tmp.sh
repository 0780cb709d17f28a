HVLA_FUSED_LN=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for v in 0 1; do HVLA_FUSED_LN=$v timeout 300 python bench.py --no-cpu-baseline --no-task-switch > gpurun_out/r2_z_b1_$v.json 2>gpurun_out/r2_z_b1_$v.err; done
