#!/bin/bash
# A/B GPU-box visit: parity suite, then the bench with each new mechanism switched off in turn (every python process under
# its own hard timeout), then an in-graph trace of the default configuration.
mkdir -p gpurun_out
TAG=${1:-ab}
timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_tests.log
run() { # name, extra args
  timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --no-cpu $2 > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
  python - <<EOF
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_$1.json").read().strip().splitlines()[-1])
    print("$1", "value %.0f" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"],
          "full %.0f" % d["e2e_full_copy"]["value"], "packed %.0f" % d["e2e_packed"]["value"], "launches", d["gpu_launches_per_step"])
except Exception as e:
    print("$1 FAILED", e)
    print(open("gpurun_out/${TAG}_bench_$1.err").read()[-1500:])
EOF
}
run default ""
run nooverlap "--overlap 0"
run torchmm "--dense-mm torch"
run c32 "--bn-act c32"
run depth1 "--e2e-depth 1"
run old "--overlap 0 --dense-mm torch --bn-act c32 --e2e-depth 1"
timeout -s KILL 300 python bench.py --trace --no-cpu > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err
timeout -s KILL 300 python bench.py --trace --no-cpu --no-pdl > gpurun_out/${TAG}_trace_nopdl.json 2> gpurun_out/${TAG}_trace_nopdl.err
ls -la gpurun_out | grep ${TAG}
