#!/usr/bin/env bash
# One 8-GPU session: real-NVLink parity at 2/4/8 processes, bench lines (prefill + train, 11B weak; config 5 inside),
# per-phase breakdown with NVLink counters.  Everything lands in gpurun_out/r2_8gpu_*.
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_tp_fused_gpu.py -m gpu -q -k "multi_process" > $O/r2_8gpu_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2_8gpu_pytest.log
timeout 400 $TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_8gpu_bench.json 2> $O/r2_8gpu_bench.err; echo "bench8 rc=$?"; tail -c 400 $O/r2_8gpu_bench.err
timeout 300 $TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --steps 10 --warmup 3 --mode train > $O/r2_8gpu_bench_train.json 2> $O/r2_8gpu_bench_train.err; echo "train8 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29703 scripts/tp_breakdown.py 11b 65536 > $O/r2_8gpu_breakdown_11b.txt 2>&1; echo "bd11 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29704 scripts/tp_breakdown.py 90b 8192 > $O/r2_8gpu_breakdown_90b.txt 2>&1; echo "bd90 rc=$?"
timeout 300 $TR --nproc-per-node 4 --master-port 29705 bench.py --gpus 4 --steps 20 --warmup 5 > $O/r2_4gpu_bench.json 2> $O/r2_4gpu_bench.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for f in ("r2_8gpu_bench.json", "r2_8gpu_bench_train.json", "r2_4gpu_bench.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "tp_parity_rel_l2")})
    print("  e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "copies_only_ms_per_step")} if d.get("e2e") else None)
    c5 = (d.get("extra") or {}).get("config5_90b_strong")
    if c5:
        print("  cfg5", {k: c5.get(k) for k in ("tokens_per_s", "ms_per_step", "efficiency_vs_1gpu_90b", "tp_parity_rel_l2")}, "train", c5.get("train"), "1gpu", c5["single_gpu"]["ms_per_step"])
PY
grep -h "nvlink\|rank 0:" $O/r2_8gpu_breakdown_11b.txt $O/r2_8gpu_breakdown_90b.txt | head -8
