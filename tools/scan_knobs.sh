#!/bin/bash
# development scan of the launch knobs (QCK_DMMA, QCK_TMA_MIN) over the benchmark workloads
for d in 0 1; do for m in 0 2048 4096 16384 1000000000; do
  echo "cz dmma=$d tma_min=$m: $(QCK_DMMA=$d QCK_TMA_MIN=$m python tools/quick_bench.py cz 10000 pade 2>&1 | grep 'F+J+H')"
done; done
for m in 0 512 2048 1000000000; do
  echo "sampling tma_min=$m: $(QCK_TMA_MIN=$m python tools/quick_bench.py sampling 200 pade 256 2>&1 | grep 'F+J+H')"
  echo "hadamard tma_min=$m: $(QCK_TMA_MIN=$m python tools/quick_bench.py hadamard 100000 pade 2>&1 | grep 'F+J+H')"
  echo "cz-exp tma_min=$m: $(QCK_TMA_MIN=$m python tools/quick_bench.py cz 10000 exponential 2>&1 | grep 'F+J+H')"
done
