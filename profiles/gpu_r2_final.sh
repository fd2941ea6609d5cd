#!/bin/bash
# round-2 final check: full GPU test suite, smoke, the driver's bench line at N=1 (default flags) and the reference arm
O=gpurun_out/r2final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err; echo "reference rc=$?"
python - <<PY
import json
j=json.loads([x for x in open("$O/bench_n1.json") if x.startswith("{")][-1])
print({k:j[k] for k in ("value","ms_per_step","gpu_launches_per_step")}, "e2e", j["e2e"]["value"], j["e2e"]["ms_per_step"], j["clocks"])
print("roofline:", j["roofline"]["kernel"][:50], j["roofline"]["us_per_launch"], j["roofline"]["frac"], j["roofline"]["traffic"])
for e in j["roofline_other"]: print("   ", e["kernel"][:70], e.get("us_per_launch"), e.get("frac"))
print("sampling", j["sampling"]["value"], j["sampling"]["ms"], "graph", j["sampling"]["graph"]["ms"], "e2e", j["sampling"]["e2e"]["ms"], "tf32", j["sampling_tf32"]["ms"])
print("cpu_baseline", j["cpu_baseline"])
r=json.loads([x for x in open("$O/bench_reference_n1.json") if x.startswith("{")][-1]); print("reference", r["value"], r["ms_per_step"])
PY
