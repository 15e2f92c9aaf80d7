python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pd in 1 0; do for g in 256 128 64; do echo "pdl $pd grid $g"; NBM_PDL=$pd python bench.py --grid $g --steps 100 --warmup 5 --no-cpu-baseline --no-flush 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['config']['cuda_graph'])
    else: print(l[:200])
"; done; done
