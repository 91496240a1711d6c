# graph replay vs eager launches, same box (run under gpurun)
python -m pytest tests -q -m gpu 2>&1 | tail -4
for cfg in "G" "E --no-graph"; do set -- $cfg; name=$1; shift; python bench.py --gpus 1 --steps 200 --warmup 20 --no-extra --no-cpu-baseline --no-e2e --profile-steps 0 "$@" > gpurun_out/abg_$name.json 2>gpurun_out/abg_$name.err || tail -5 gpurun_out/abg_$name.err; done
for cfg in "G2" "E2 --no-graph"; do set -- $cfg; name=$1; shift; python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e --profile-steps 0 "$@" > gpurun_out/abg_$name.json 2>gpurun_out/abg_$name.err || tail -5 gpurun_out/abg_$name.err; done
python - <<PY
import json
for f in ("G","E","G2","E2"):
    try:
        j=json.load(open("gpurun_out/abg_%s.json"%f))
        print(f, round(j["value"]), round(j["ms_per_step"],4), j["launch_mode"][:20], j["step_times"])
    except Exception as e: print(f, "ERR", e)
PY
