mkdir -p gpurun_out
S=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $S/r2_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-c5 > $S/b_launch.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 > $S/r2_bench_final_n1.json 2> $S/r2_bench_final_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $S/r2_bench_final_ref.json 2>&1
tail -c 1500 $S/r2_bench_final_n1.json; echo; tail -c 600 $S/r2_bench_final_ref.json; ls -la $S | tail -8
