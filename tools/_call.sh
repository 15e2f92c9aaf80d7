mkdir -p gpurun_out/r2_named
python bench.py --gpus 1 --steps 30 --warmup 3 --workload poisson_boltzmann --no-cpu-baseline > gpurun_out/r2_named/pb256_1gpu.json 2> gpurun_out/r2_named/pb256_1gpu.err
bash tools/run_named.sh 2
