#!/bin/bash
# one-call set-up (polee_set_sample) with / without the kernel preload thread against the three-call sequence:
# warm (alternating in one process) and one-shot (fresh processes), with the library's phase timing
mkdir -p gpurun_out
O=gpurun_out
T=${1:-e2}
python -m pytest tests -m gpu -q > $O/r02_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest.log
POLEE_SETUP_TIMING=1 python tools/e2e_breakdown.py --reps 4 > $O/r02_${T}_e2e_alt.log 2>&1
for i in 1 2; do
  POLEE_SETUP_TIMING=1 POLEE_SET_SAMPLE=3calls python bench.py --oneshot-child > $O/r02_${T}_oneshot_3calls_$i.log 2>&1
  POLEE_SETUP_TIMING=1 POLEE_PRELOAD=0 python bench.py --oneshot-child > $O/r02_${T}_oneshot_onecall_nopreload_$i.log 2>&1
  POLEE_SETUP_TIMING=1 python bench.py --oneshot-child > $O/r02_${T}_oneshot_onecall_$i.log 2>&1
done
tail -3 $O/r02_${T}_pytest.log | cut -c1-300; grep total $O/r02_${T}_e2e_alt.log | cut -c1-300; grep -H oneshot_fit $O/r02_${T}_oneshot_*.log
