#!/usr/bin/env bash
# Per-kernel device times of one training step (ncu launch list; serialised, cold cache): fwd gate/up(+caches), fwd down,
# d_act + SiLU' epilogue, two-phase dX, three wgrads.  Usage: bash scripts/train_launches.sh [tag]
TAG=${1:-tmp}
mkdir -p gpurun_out
python scripts/profile_targets.py train > /dev/null 2>&1 || { echo "plain run failed"; exit 1; }
ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"gemm_kernel" -c 21 --csv \
  --log-file gpurun_out/${TAG}_launches_train.csv python scripts/profile_targets.py train > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches_train.csv")) if len(r)>5 and r[0].isdigit()]
names=["fwd gate/up+caches","fwd down","d_act+SiLU'","dX two-phase","wgrad gate","wgrad up","wgrad down"]
fl=[4,2,2,4,2,2,2]
base=2*8192*4096*14336/1e12
tot=0
for n,f,r in zip(names,fl,rows[-7:]):
    us=float(r[-1])/1e3; tot+=us
    print(f"{n:20s} {us:8.1f} us  {f/2*base/us*1e6:7.0f} TFLOP/s  {r[4][28:52]}")
print(f"sum {tot:.1f} us")
PY
