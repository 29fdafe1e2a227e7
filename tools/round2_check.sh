#!/bin/bash
# one GPU call: the whole GPU suite, smoke(), and the default bench line (what the driver runs at round end)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -3 gpurun_out/r2_pytest.log
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3), "roof", round(d["roofline"]["frac"], 3),
      {k: round(v["ms_per_step"] * 1000, 1) for k, v in d["kernels"].items()},
      "c3", round(d.get("config3", {}).get("ms_per_step", 0), 4), "c4", round(d.get("config4", {}).get("ms_per_step", 0), 4), d["gpu_launches"], d["clocks"], d.get("check"))
P
# the launch list of the same bench command (cold, serialised: shares only), after the clean run above
if [ "$1" = "launches" ]; then
  timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
  echo "ncu rc=$?"
fi
