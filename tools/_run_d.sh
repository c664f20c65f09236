python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_d.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_d.log
ARGS="--steps 2 --warmup 3 --no-north-star --no-shim --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/r02_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:raycast2_kernel|score_blocks' -s 36 -c 3 -f -o gpurun_out/r02_prof_c2 python bench.py $ARGS > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py $ARGS > gpurun_out/r02_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_e.err
ls -la gpurun_out | tail -8
