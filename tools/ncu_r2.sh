# Round-2 ncu evidence (run under gpurun, ONE GPU): launch list of 5 eager steps + one `--set full` capture of a whole step.
# Each ncu command runs only after the identical plain command exited 0.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra --profile-steps 0 --no-graph --no-prefetch"
KRE='regex:^(k1_|layer0_|spmm_|gemm_3xtf32|readout_|ln_relu|loss_|colsum_|bn_|adamw_)'
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/r2_launches_ncu.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
echo LIST_RC=$?
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --metrics lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none --import-source on -k "$KRE" -s 84 -c 28 -f -o gpurun_out/r2_step_full $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo FULL_RC=$?
ls -la gpurun_out/ | grep r2_
tail -3 gpurun_out/r2_ncu_full.log
