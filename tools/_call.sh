python examples/solve_named.py stars --epochs 40 2>&1 | tail -2
python examples/solve_named.py dragon_like --epochs 40 2>&1 | tail -2
python examples/solve_named.py poisson_boltzmann --epochs 20 2>&1 | tail -2
python examples/solve_named.py poisson_boltzmann --epochs 8 --n-train 64 --preconditioner 2>&1 | tail -2
python examples/solve_named.py dragon_like --epochs 20 --multi-gpu 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 examples/solve_named.py poisson_boltzmann --epochs 20 --multi-gpu 2>&1 | grep -v "Warning\|OMP\|\*\*\*" | tail -2
python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench 1gpu', d['ms_per_step'], d['value'], d['e2e']['value'])"
