timeout 600 python -m pytest tests/test_celt_synth.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-cfg4 --no-sustain 2>/dev/null > gpurun_out/glue.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/glue.json').read().strip().splitlines()[-1]); c=d["decode_chain"]["celt_entropy"]; print(d["value"], c["frames"], c["full_decode"]["frames"], c["full_decode"]["ms"])
PY
