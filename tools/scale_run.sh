set -x
P=${PREFIX:-s8}   # file prefix under gpurun_out/
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/${P}_n1.json 2> gpurun_out/${P}_n1.err
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --profile-steps 0 > gpurun_out/${P}_n$n.json 2> gpurun_out/${P}_n$n.err; echo RC$n=$?
done
$TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 200 --warmup 20 --profile-steps 0 --no-e2e > gpurun_out/${P}_n8_200.json 2> gpurun_out/${P}_n8_200.err; echo RC8L=$?
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --profile-steps 0 --no-e2e --no-graph --no-variants > gpurun_out/${P}_n8_eager.json 2> gpurun_out/${P}_n8_eager.err; echo RC8E=$?
$TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --workload wide --steps 20 --warmup 5 --profile-steps 0 --no-e2e > gpurun_out/${P}_wide_n8.json 2> gpurun_out/${P}_wide_n8.err; echo RCW8=$?
python bench.py --gpus 1 --workload wide --steps 20 --warmup 5 --profile-steps 0 --no-e2e --no-extra --no-cpu-baseline > gpurun_out/${P}_wide_n1.json 2> gpurun_out/${P}_wide_n1.err
$TR --nproc-per-node 8 --master-port 29523 tools/dp_parity.py --branch multicast --out gpurun_out/dp_parity_n8.json 2>&1 | tail -3
PREFIX=$P python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/%s_*.json" % __import__("os").environ.get("PREFIX", "s8"))):
    try:
        j=json.load(open(f))
        print(f, "value", round(j["value"]), "ms", round(j["ms_per_step"],4), j.get("launch_mode","")[:12], "clk", j["clocks"], "e2e", j["e2e"] and j["e2e"].get("value"))
        print("    step_times", j.get("step_times"))
        print("    variants", j.get("sampling_variants"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 600 gpurun_out/${P}_n8.err
