#!/usr/bin/env bash
# Round-2 evidence capture on ONE B200 (run through gpurun): every profile names the digest of the libl32ffn.so it saw.
#   1. launch list of the bench command (gpu__time_duration per kernel; kernel SHARES of the step)
#   2. ncu --set full of the shipped kernels: fused gate/up + down (prefill), the backward GEMMs (d_act + SiLU', two-phase
#      dX, wgrad), small-M decode at B = 64, Add-RMSNorm forward / backward
# Every target is first run WITHOUT ncu and must exit 0.  Raw .ncu-rep files land in gpurun_out/ (scratch); the
# summaries under profiles/ are made on the CPU box by scripts/ncu_summary.py.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TAG=${TAG:-r2}
LIB=llama-3.2-multimodal_b200/libl32ffn.so
DIGEST=$(sha256sum $LIB | cut -c1-16)
# the .so hash changes from build to build of the SAME sources (nvcc's anonymous-namespace mangling); the source digest
# (llama-3.2-multimodal_b200/build.py: sha256 over csrc/*.cu, the headers and the flags) is the reproducible one
SRC=$(cut -c1-16 llama-3.2-multimodal_b200/_build/digest.txt 2>/dev/null)
echo "src_sha256_16=$SRC" > $OUT/${TAG}_src_digest.txt
echo "lib_sha256_16=$DIGEST  src_sha256_16=$SRC  $(date -u +%FT%TZ)  $(nvidia-smi --query-gpu=name,driver_version --format=csv,noheader)" > $OUT/${TAG}_digest.txt
NCU="ncu --clock-control none"
full() {   # name, kernel regex, skip, count, target args...
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  if timeout 300 python scripts/profile_targets.py "$@" > $OUT/${TAG}_plain_$name.log 2>&1; then
    timeout 900 $NCU --set full --import-source on -k regex:$regex --launch-skip $skip -c $count -f -o $OUT/${TAG}_full_$name \
      python scripts/profile_targets.py "$@" > $OUT/${TAG}_ncu_$name.log 2>&1
    echo "$name: ncu rc=$?" >> $OUT/${TAG}_digest.txt
    # summarise on the box (gpurun_out/ may carry at most 64 MiB back) and drop the raw report unless asked to keep it
    python scripts/ncu_summary.py $OUT/${TAG}_full_$name.ncu-rep "lib_sha256_16=$DIGEST src_sha256_16=$SRC; ncu --set full --clock-control none --import-source on; python scripts/profile_targets.py $*; $(date -u +%F)" > $OUT/${TAG}_ncu_full_$name.txt 2>/dev/null
    if [ "${KEEP_REP:-0}" = 0 ]; then rm -f $OUT/${TAG}_full_$name.ncu-rep; fi
  else
    echo "$name: plain run FAILED, not profiled" >> $OUT/${TAG}_digest.txt
  fi
}
if [ "${SKIP_LAUNCHES:-0}" = 0 ]; then
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 80 --csv --log-file $OUT/${TAG}_launches_bench_11b.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > $OUT/${TAG}_ncu_launches.log 2>&1
  timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:"gemm_kernel|rmsnorm" -c 40 --csv --log-file $OUT/${TAG}_launches_train_11b.csv \
    python scripts/profile_targets.py train > $OUT/${TAG}_ncu_launches_train.log 2>&1
fi
for t in ${TARGETS:-prefill train decode norm}; do
  case $t in
    prefill) full prefill gemm_kernel 4 2 prefill ;;
    train)   full train gemm_kernel 7 7 train ;;
    decode)  full decode_b64 ffn_decode 14 2 decode ;;
    norm)    full norm rmsnorm 8 4 norm ;;
    attention) full attention "gqa_attention|rope_kv" 2 5 attention ;;
    lmhead)  full lmhead "gemm_kernel|ce_" 3 3 lmhead ;;
  esac
done
ls -la $OUT/${TAG}_* >> $OUT/${TAG}_digest.txt
cat $OUT/${TAG}_digest.txt
