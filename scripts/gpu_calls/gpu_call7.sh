#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/c7_pytest.log 2>&1
tail -3 gpurun_out/c7_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/c7_smoke.log 2>&1; tail -1 gpurun_out/c7_smoke.log
for m in resnet50 wrn alexnet; do
  timeout 200 python bench.py --model $m --steps 20 --warmup 5 > gpurun_out/c7_bench_$m.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/c7_bench_$m.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("$m", d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz"), d.get("final_loss"))
else:
    print("$m FAILED"); print(open("gpurun_out/c7_bench_$m.log").read()[-1500:])
PY
done
