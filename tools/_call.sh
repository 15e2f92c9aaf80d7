python -m pytest tests/test_gpu_multidevice.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2_multidevice_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/r2_check_multigpu_2.log
for b in 1 0; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$b bench.py --gpus 2 --steps 50 --warmup 5 --no-weak --balance $b > gpurun_out/bench_r2o_2gpu_bal$b.json 2> gpurun_out/bench_r2o_2gpu_bal$b.err
tail -2 gpurun_out/bench_r2o_2gpu_bal$b.err
done
