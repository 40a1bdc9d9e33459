mkdir -p gpurun_out
# every launch of one bench step (warm-up launches skipped), device time per launch
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/b_launch.log 2>&1
# full captures of the three product kernels
ncu --set full --clock-control none --import-source on -k regex:k2_dispersion_fast -c 1 -o gpurun_out/prof_k2_final python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_column -c 1 -o gpurun_out/prof_k1_final python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_coopw -c 1 -o gpurun_out/prof_k2coop_final python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/b3.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2>&1
tail -c 400 gpurun_out/bench_final_n1.json; ls -la gpurun_out | tail -8
