# Round-2 final ncu evidence (run under gpurun, ONE GPU).  Each ncu command runs only after the identical plain command exited 0.
# (1) launch list of 5 eager training steps, (2) `--set full` capture of one whole training step (28 launches),
# (3) `--set full` capture of the planes GEMM (gemm_planes_kernel<2>) inside batch-4096 inference.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra --profile-steps 0 --no-graph --no-prefetch"
KRE='regex:^(k1_|layer0_|spmm_|gemm_3xtf32|gemm_planes|split_planes|readout_|ln_relu|loss_|colsum_|bn_|adamw_)'
$CMD > gpurun_out/r2b_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/r2b_launches_ncu.csv $CMD > gpurun_out/r2b_ncu_list.log 2>&1
echo LIST_RC=$?
$CMD > gpurun_out/r2b_plain2.log 2>&1 &&
ncu --set full --metrics lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none --import-source on -k "$KRE" -s 84 -c 28 -f -o gpurun_out/r2b_step_full $CMD > gpurun_out/r2b_ncu_full.log 2>&1
echo FULL_RC=$?
ICMD="python bench.py --workload infer --molecules 32768 --steps 4 --warmup 2 --no-cpu-baseline --no-e2e --profile-steps 0"
$ICMD > gpurun_out/r2b_infer_plain.log 2>&1 &&
ncu --set full --metrics lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none --import-source on -k "regex:^(gemm_planes|spmm_mol|split_planes)" -s 6 -c 6 -f -o gpurun_out/r2b_infer_planes $ICMD > gpurun_out/r2b_ncu_infer.log 2>&1
echo INFER_RC=$?
ls -la gpurun_out/ | grep r2b_
tail -2 gpurun_out/r2b_ncu_full.log gpurun_out/r2b_ncu_infer.log
