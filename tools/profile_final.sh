# tools/profile_final.sh -- the round's ncu evidence (run under gpurun, one GPU).  The .ncu-rep files are summarised ON
# THE BOX by tools/ncu_summary.py (raw page + SASS opcode mix) and deleted: gpurun brings back at most 64 MiB.
mkdir -p gpurun_out
S=gpurun_out
cap() { # cap <kernel regex> <skip> <out name> <note> <command...>
  local k="$1" s="$2" o="$3" note="$4"; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s "$s" -c 1 -o $S/$o "$@" > $S/$o.log 2>&1
  python tools/ncu_summary.py $S/$o.ncu-rep $S/$o.txt "$note" >> $S/$o.log 2>&1
  rm -f $S/$o.ncu-rep
}
# every launch of one bench step (warm-up launches included in the list; bench default workload, no C5 block)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $S/r2_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-c5 > $S/b_launch.log 2>&1
cap "k2_(dispersion|coop2)" 1 r2_k2_final "Round 2, K2 dispersion kernel of the bench step (C2 x 32 models)" python bench.py --steps 1 --warmup 1 --no-cpu --no-c5
cap k1_box 1 r2_k1_final "Round 2, K1 nearest-nucleus kernel (k1_box_kernel), C2 x 32 models in one launch" python bench.py --steps 1 --warmup 1 --no-cpu --no-c5
cap k2_coopw 0 r2_k2coop_final "Round 2, proposal-sized dispersion call (several warps per column)" python bench.py --steps 1 --warmup 1 --no-cpu --no-c5
cap k1_box 3 r2_k1_c5 "Round 2, K1 (k1_box_kernel) on C5: 1024x1024x80 nodes, 5000 nuclei" python tools/k1_sweep.py C5 auto
cap grt_kernel 1 r2_grt "Round 2, generalized R/T kernel (grt_kernel): 1024 low-velocity columns, Rayleigh, 11 frequencies" python tools/grt_bench.py 32 1
cap fm2d_kernel 1 r2_fm2d "Round 2, fast-marching kernel (fm2d_kernel): example1 geometry, 88 problems of 101x101 nodes" python tools/fm2d_bench.py
python tools/fm2d_bench.py > $S/r2_fm2d.json 2>$S/fm2d.err
python tools/likelihood_bench.py > $S/r2_likelihood_step.json 2>$S/like.err
python tools/grt_bench.py 64 1 > $S/r2_grt_rayleigh.json 2>$S/grt1.err
python tools/grt_bench.py 64 0 > $S/r2_grt_love.json 2>$S/grt0.err
for c in C2x32 C5 C3 C1; do python tools/k1_sweep.py $c auto; done > $S/r2_k1_sweep.log 2>&1
python bench.py --steps 10 --warmup 3 > $S/r2_bench_final_n1.json 2> $S/r2_bench_final_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $S/r2_bench_final_ref.json 2>&1
tail -c 300 $S/r2_bench_final_n1.json; ls -la $S | tail -12
