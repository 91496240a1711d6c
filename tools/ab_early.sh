python -m pytest tests -q -m gpu 2>&1 | tail -4
for cfg in "Y EIMS_EARLY_DIMS=1" "N EIMS_EARLY_DIMS=0" "Y2 EIMS_EARLY_DIMS=1" "N2 EIMS_EARLY_DIMS=0"; do set -- $cfg; name=$1; shift; env "$@" python bench.py --gpus 1 --steps 200 --warmup 20 --no-extra --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/abe_$name.json 2>gpurun_out/abe_$name.err || tail -5 gpurun_out/abe_$name.err; done
python - <<PY
import json
for f in ("Y","N","Y2","N2"):
    try:
        j=json.load(open("gpurun_out/abe_%s.json"%f))
        print(f, round(j["value"]), round(j["ms_per_step"],4), j["step_times"])
    except Exception as e: print(f, "ERR", e)
PY
