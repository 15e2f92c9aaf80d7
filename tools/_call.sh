python -m pytest tests/test_gpu_points.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -6
for z in 1 2; do python bench.py --grid 128 --zoom $z --steps 30 | tail -1; done
NBM_ZOOM1_SHARED=0 python bench.py --grid 128 --zoom 1 --steps 30 | tail -1
