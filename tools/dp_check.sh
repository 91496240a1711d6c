# data-parallel parity on 2 GPUs (run under gpurun --gpus 2)
python -m pytest tests/test_gpu_dp.py -q -m gpu 2>&1 | tail -6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29541 tools/dp_parity.py --branch multicast --out gpurun_out/dp_parity_multicast.json 2>&1 | grep -E "DP PARITY|Error|error" | head -3
$TR --master-port 29542 tools/dp_parity.py --branch peer --out gpurun_out/dp_parity_peer.json 2>&1 | grep -E "DP PARITY|Error|error" | head -3
$TR --master-port 29543 tools/dp_parity.py --graph --out gpurun_out/dp_parity_graph.json 2>&1 | grep -E "DP PARITY|Error|error" | head -3
python - <<PY
import json
for b in ("multicast","peer","graph"):
    try:
        j=json.load(open("gpurun_out/dp_parity_%s.json"%b))
        print(b, {k:j[k] for k in ("branch","graph","kernel_vs_nccl_plus_adamw","ranks_bit_identical","lost_peer","loss_rel_err_max","param_abs_err_over_bound_max","reliable_fraction_mean")})
    except Exception as e: print(b,"ERR",e)
PY
