#!/bin/bash
# ncu --set full of the MPPI step for one kernel plan:  tools/prof_ws.sh <tag> <MPCB_MPPI_WS value> [K H]
tag=$1; ws=$2; K=${3:-65536}; H=${4:-100}
export MPCB_MPPI_WS=$ws
python tools/prof_mppi.py $K $H f32 6 > gpurun_out/prof_${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mppi -s 3 -c 2 -o gpurun_out/prof_${tag} -f python tools/prof_mppi.py $K $H f32 6 > gpurun_out/prof_${tag}_ncu.log 2>&1
echo "prof rc=$?"
