#!/usr/bin/env bash
# One process per case (a trapped kernel kills only its own CUDA context); everything lands in gpurun_out/probe.log
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
: > $LOG
run() { echo "=== $*" >> $LOG; timeout 120 python scripts/gpu_probe.py "$@" >> $LOG 2>&1; echo "exit=$?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run rmsnorm
run gemm 128 256 64 1
run gemm 128 256 256 1
run gemm 256 512 512 1
run gemm 300 520 328 1
run gemm 256 256 64 2
run gemm 256 256 512 2
run gemm 1024 1024 1024 2
run gemm 520 776 328 2
run gemm16 512 512 512 2
run gemm 256 512 256 1 0 1
run gemm 256 512 256 1 1 0
run gemm 256 512 256 1 1 1
run gemm 512 512 512 2 0 1
run gemm 512 512 512 2 1 1
run gemm 512 512 512 2 0 1 1
run swiglu 256 256 688 0
run swiglu 256 256 688 1
run swiglu 100 256 688 1
run swiglu 1024 4096 14336 1
run ffn_bwd 256 256 688
run ffn_bwd 1024 1024 2048
run perf 4096 14336 8192
run perf 8192 28672 8192
tail -n 400 $LOG
