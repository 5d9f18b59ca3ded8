# round-2 profiling commands (B200_PROFILING.md recipe); run as  gpurun -- bash scripts/profile_round2.sh A|B
set -x
if [ "$1" = "A" ]; then
E="python bench.py --steps 2 --warmup 3 --no-cpu --no-solves --no-colloc --no-single"
$E > gpurun_out/plain_eval.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_eval.csv $E > gpurun_out/ncu1.log 2>&1
$E > gpurun_out/plain_eval.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rk4_ -s 20 -c 2 -o gpurun_out/r02_rk4 $E > gpurun_out/ncu2.log 2>&1
K="python tests/gpu_kkt_bench.py race_param_rk4_drone 444"
$K > gpurun_out/plain_kkt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kkt_factor_kernel\|kkt_solve_kernel -s 2 -c 2 -o gpurun_out/r02_kkt $K > gpurun_out/ncu3.log 2>&1
elif [ "$1" = "B" ]; then
K2="python tests/gpu_kkt_bench.py fig8_global_colloc_drone 2"
$K2 > gpurun_out/plain_kktc.log 2>&1 && ncu --set full --clock-control none -k regex:kktc_ -s 3 -c 3 -o gpurun_out/r02_kktc $K2 > gpurun_out/ncu4.log 2>&1
C="python tests/gpu_eval_bench.py 256 fig8_global_colloc_drone"
$C > gpurun_out/plain_colloc.log 2>&1 && ncu --set full --clock-control none -k regex:colloc -s 8 -c 2 -o gpurun_out/r02_colloc $C > gpurun_out/ncu5.log 2>&1
S="python tests/gpu_solves_profile.py 64 64 4"
$S > gpurun_out/plain_solves.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 6000 --csv --log-file gpurun_out/r02_launches_solves.csv $S > gpurun_out/ncu6.log 2>&1
fi
ls -la gpurun_out/
# tail tape kernel (open racelines): bash scripts/profile_round2.sh T
if [ "$1" = "T" ]; then
T="python tests/gpu_tail_time.py"
$T > gpurun_out/plain_tail.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tail_tape -s 2 -c 1 -o gpurun_out/r02_tail $T > gpurun_out/ncu7.log 2>&1
fi
