# round-2 standard GPU check: full GPU test suite, smoke, bench (N=1), optional A/B against the cooperative chains
O=gpurun_out/${1:-r2x}; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest.log; cat $O/pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 400 $O/bench.err
python - <<PY
import json
j=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print({k:j[k] for k in ("value","ms_per_step","gpu_launches_per_step")}, j["e2e"]["value"], j["e2e"]["ms_per_step"])
print("roofline:", j["roofline"]["kernel"][:40], j["roofline"]["us_per_launch"], j["roofline"]["frac"])
for e in j["roofline_other"]:
    print("   ", e["kernel"][:60], e.get("us_per_launch"), e.get("frac"))
print(j["sampling"])
PY
