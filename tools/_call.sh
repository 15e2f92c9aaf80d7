K='regex:^(void )?(nbm::)?(fwd_nodes|points_|cube_|node_grad|reduce_partials|prep_params|precond)'
for z in 1 2; do
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/launches_r2ad_zoom$z.csv python bench.py --grid 128 --zoom $z --steps 3 --warmup 3 > gpurun_out/ncu_bench_r2ad_zoom$z.log 2>&1
done
