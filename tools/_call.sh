mkdir -p gpurun_out
S=gpurun_out/r2_sanitizer.txt
echo "compute-sanitizer on 1 x B200, round 2 (TMA stencil kernel, list chain on the side stream + merge, balanced runs of the network kernels, fused tail kernels)" > $S
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_shared.py -q -x -k "list_chain or stencil_tma_is_bitwise or rows_loss_and_gradient" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|error" | head -8 | sed 's/^/  memcheck : /' >> $S
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_shared.py -q -x -k "list_chain and (sphere-16 or star-32) or stencil_tma_is_bitwise and (sphere-16-32 or star-15)" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|hazard" | head -8 | sed 's/^/  racecheck: /' >> $S
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_points.py -q -x -k "general_path or region_scaled" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head -6 | sed 's/^/  memcheck (general path): /' >> $S
cat $S
