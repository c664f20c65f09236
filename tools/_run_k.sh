python -m pytest tests/test_guards.py tests/test_mesh_normals.py -m gpu -x -q 2>&1 | tail -3
VARIANTS=6,7 python tools/bench_variants.py C2 3 2>&1 | tail -2
VARIANTS=6,7 python tools/bench_variants.py C3 3 2>&1 | tail -2
ARGS="--steps 2 --warmup 3 --no-north-star --no-shim --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/r02_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:raycast2_kernel|score_blocks' -s 36 -c 3 -f -o gpurun_out/r02_prof_c2 python bench.py $ARGS > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py $ARGS > gpurun_out/r02_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
python bench.py --normals mesh --steps 5 --warmup 3 --no-north-star 2>/dev/null > gpurun_out/r02_bench_c2_mesh.json; echo "mesh bench rc=$?"
