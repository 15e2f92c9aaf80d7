python -m pytest tests/test_gpu_points.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4
for ck in 1 0; do echo "cube kernels $ck"; for z in 1 2 3; do NBM_CUBE_KERNELS=$ck python bench.py --grid 128 --zoom $z --steps 30 | tail -1 | cut -c1-125; done; done
