# A/B of the aggregation kernels on one GPU (bench stage timings); run under gpurun
for cfg in "M EIMS_SPMM_MOL=1" "C EIMS_SPMM_MOL=0"; do set -- $cfg; name=$1; shift; env "$@" python bench.py --gpus 1 --workload infer --molecules 262144 > gpurun_out/abi_$name.json 2>gpurun_out/abi_$name.err || tail -5 gpurun_out/abi_$name.err; done
python - <<PY
import json
for f in "MC":
    try:
        j=json.load(open("gpurun_out/abi_%s.json"%f))
        st=j["stages"]
        print(f, round(j["value"]), round(j["ms_per_step"],4), {k: (st[k]["ms_per_step"], st[k]["frac"]) for k in ("k1_batch_build","spmm_fwd","readout","gemm_gcn_fwd")})
    except Exception as e: print(f, "ERR", e)
PY
