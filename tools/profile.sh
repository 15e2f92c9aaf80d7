#!/bin/bash
# Round profile of the step on ONE B200 (run under gpurun): launch list + one `ncu --set full` capture of the dense
# step kernels.  Outputs go to gpurun_out/; summaries are copied into profiles/ by tools/summarise_profile.py.
set -u
TAG=${1:-r2a}
shift || true
EXTRA="$*"
K='regex:^(void )?(nbm::)?(stencil_tma::)?(fwd_nodes|residual|adjoint|stencil_tma|node_grad|extrap|irregular|reduce_partials|apply_update|prep_params|precond|finalize_step|step_)'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph $EXTRA \
    > gpurun_out/ncu_bench_${TAG}.log 2>&1
# skip the set-up and warm-up launches: capture the dense step kernels of a late step
timeout 500 ncu --set full --import-source on --clock-control none \
    -k 'regex:(fwd_nodes_kernel|stencil_tma_kernel|node_grad_kernel|merge_lists_kernel|irregular_fb_kernel|finalize_step_kernel)' --launch-skip 12 -c 6 \
    -f -o gpurun_out/prof_${TAG} python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph $EXTRA \
    > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
timeout 200 python bench.py --steps 100 --warmup 5 $EXTRA > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err
tail -c 400 gpurun_out/bench_${TAG}_1gpu.json
