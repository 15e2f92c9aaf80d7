python -m pytest tests/test_gpu_points.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -6
for f in 1 0; do echo "fused points $f"; for z in 2 3; do NBM_POINTS_FUSED=$f python bench.py --grid 128 --zoom $z --steps 30 | tail -1 | cut -c1-125; done; done
