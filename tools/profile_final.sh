# tools/profile_final.sh -- the round's ncu evidence (run under gpurun, one GPU); outputs land in gpurun_out/ and are
# summarised into profiles/ by tools/ncu_summary.py + tools/k2_traffic.py here.
mkdir -p gpurun_out
# every launch of one bench step (warm-up launches included in the list; bench default workload, no C5 block)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-c5 > gpurun_out/b_launch.log 2>&1
# full captures: the dispersion kernel of the bench step (whatever shape launch_k2 picked), the nearest-nucleus kernel,
# and the proposal-sized cooperative kernel
ncu --set full --clock-control none --import-source on -k regex:"k2_(dispersion|coop2)" -s 1 -c 1 -o gpurun_out/r2_prof_k2_final python bench.py --steps 1 --warmup 1 --no-cpu --no-c5 > gpurun_out/b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_tile -s 1 -c 1 -o gpurun_out/r2_prof_k1_final python bench.py --steps 1 --warmup 1 --no-cpu --no-c5 > gpurun_out/b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_coopw -c 1 -o gpurun_out/r2_prof_k2coop_final python bench.py --steps 1 --warmup 1 --no-cpu --no-c5 > gpurun_out/b3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_tile -s 3 -c 1 -o gpurun_out/r2_prof_k1_c5 python tools/k1_bench.py C5 > gpurun_out/b4.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2>&1
tail -c 300 gpurun_out/r2_bench_final_n1.json; ls -la gpurun_out | tail -8
