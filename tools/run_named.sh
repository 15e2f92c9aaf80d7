#!/bin/bash
# BASELINE.json's named configurations at N GPUs of one box (run under gpurun --gpus N): one JSON line each under
# gpurun_out/r2_named/.   usage: tools/run_named.sh N
set -u
N=${1:-1}
OUT=gpurun_out/r2_named
mkdir -p $OUT
port=29600
run() {   # name, bench args...
    local name=$1; shift
    port=$((port + 1))
    if [ "$N" = "1" ]; then
        timeout 600 python bench.py --gpus 1 --steps 30 --warmup 3 "$@" > $OUT/${name}_${N}gpu.json 2> $OUT/${name}_${N}gpu.err
    else
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
            bench.py --gpus $N --steps 30 --warmup 3 "$@" > $OUT/${name}_${N}gpu.json 2> $OUT/${name}_${N}gpu.err
    fi
    echo "$name N=$N rc=$? $(tail -c 200 $OUT/${name}_${N}gpu.json | tr '\n' ' ')"
}
if [ "$N" != "1" ]; then
    port=$((port + 1))
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
        tools/check_multigpu.py 2>&1 | grep -v Warning | tail -3 > $OUT/check_multigpu_${N}gpu.log
    cat $OUT/check_multigpu_${N}gpu.log
    run sphere256 --no-cpu-baseline
    run sphere256_equal_slabs --no-cpu-baseline --no-weak --balance 0
    run pb256 --workload poisson_boltzmann --no-cpu-baseline --no-weak
else
    run stars64 --workload stars --grid 64 --lvl 128 --no-cpu-baseline
    run pb256 --workload poisson_boltzmann --no-cpu-baseline
fi
if [ "$N" = "1" ] || [ "$N" = "8" ]; then
    run dragon128 --workload dragon_like --grid 128 --interp quadratic --no-cpu-baseline --no-weak
fi
for g in 64 128 512; do
    run sphere$g --grid $g --no-cpu-baseline --no-weak
done
