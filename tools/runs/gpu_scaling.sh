#!/bin/bash
# The driver's scaling run, by hand: bench.py at N = 1, 2, 4, 8 (one torchrun per N, each under its own timeout),
# lines collected in gpurun_out/scale_<tag>.jsonl.   gpurun --gpus 8 -- 'bash tools/gpu_scaling.sh r02'
TAG=${1:-manual}
mkdir -p gpurun_out
OUT=gpurun_out/scale_${TAG}.jsonl
: > "$OUT"
timeout -k 5 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/scale_${TAG}_n1.err | tail -n 1 >> "$OUT"
for N in 2 4 8; do
  NG=$(python -c "import torch; print(torch.cuda.device_count())")
  [ "$NG" -lt "$N" ] && break
  timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/scale_${TAG}_n${N}.err | tail -n 1 >> "$OUT"
done
python - "$OUT" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    ln = ln.strip()
    if ln.startswith("{"):
        d = json.loads(ln)
        print(d["n_gpus"], "GPU(s):", round(d["value"], 2), d["unit"], "| ms/epoch", round(d["ms_per_step"], 2), "| e2e", round(d["e2e"]["value"], 2),
              "| comm", round(1e3 * d.get("comm_s_per_epoch", 0.0), 3), "ms | SpMM share", round(d["roofline"]["share_of_step"], 3))
PY
