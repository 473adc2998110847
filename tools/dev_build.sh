#!/bin/bash
# Developer helper: rebuild the kernel-logic emulator, run its tests, rebuild libscpb200.so and (with "timers") the
# phase-timer variant.
set -e
cd /root/repo
python -c "
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from emu import emu
emu.build(True)"
timeout 3000 python -m pytest tests/test_kernel_logic_emu.py -x -q 2>&1 | tail -2
P=/root/repo/senquential-convex-programming-for-trajectory-planning_b200
python $P/build.py --force -v 2>&1 | grep -A3 "k_scp_solveILb1ELi8" | grep -E "registers|spill" || true
if [ "$1" = "timers" ]; then python $P/build.py --force -DSCP_PHASE_TIMERS $P/libvariant_timers.so; fi
echo build-ok
