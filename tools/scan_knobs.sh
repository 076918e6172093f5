#!/bin/bash
# development scan of the launch knobs over the benchmark workloads (each knob is read once per process)
for env in "" "QCK_ROWSLICE=0" "QCK_ROWSLICE=0 QCK_DMMA=1" "QCK_ROWSLICE_DENSE=1" "QCK_ROWSLICE_WARPS=6" "QCK_ROWSLICE_WARPS=7"; do
  echo "cz [$env]: $(env $env python tools/quick_bench.py cz 10000 pade 2>&1 | grep 'F+J+H')"
done
for env in "" "QCK_COLUMN=0"; do
  echo "sampling [$env]: $(env $env python tools/quick_bench.py sampling 200 pade 256 2>&1 | grep 'F+J+H')"
  echo "hadamard [$env]: $(env $env python tools/quick_bench.py hadamard 100000 pade 2>&1 | grep 'F+J+H')"
  echo "ket [$env]: $(env $env python tools/quick_bench.py ket 100000 pade 2>&1 | grep 'F+J+H')"
done
echo "cz-exp: $(python tools/quick_bench.py cz 10000 exponential 2>&1 | grep 'F+J+H')"
