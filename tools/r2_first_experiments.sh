#!/bin/bash
# First GPU call of the next round: the two predication experiments left compiled-but-unmeasured at the end of round 1
# (worklist item 1).  Build here first:   VARIANTS="$(sed -n 's/^# VARIANTS=//p' tools/r2_first_experiments.sh)" bash tools/build_variants.sh
# VARIANTS=BASE:-DMGS_FWD_PREDICATED=0;PF:-DMGS_FWD_PREDICATED=1;PB:-DMGS_BWD_PREDICATED=1;PFB:-DMGS_FWD_PREDICATED=1 -DMGS_BWD_PREDICATED=1
export VARIANTS="BASE:-DMGS_FWD_PREDICATED=0;PF:-DMGS_FWD_PREDICATED=1;PB:-DMGS_BWD_PREDICATED=1;PFB:-DMGS_FWD_PREDICATED=1 -DMGS_BWD_PREDICATED=1"
PARITY=1 WORKLOADS="c3 c2 mg" bash tools/gpu_variants.sh
