#!/bin/bash
# The ncu evidence of a round, in one gpurun call (every ncu run follows a plain run of the same command that exited 0):
#   1. launch list of the bench command (per-launch device time; cold-cache and serialised: compare SHARES)
#   2. --set full of the MPPI step kernel on configs[1]
#   3. --set full of the fused UKF kernel on configs[2]
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/prof_${tag}_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${tag}_bench.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/prof_${tag}_bench_ncu.log 2>&1
echo "launch list rc=$?"
python tools/prof_mppi.py 65536 100 f32 6 > gpurun_out/prof_${tag}_mppi_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mppi -s 3 -c 2 -o gpurun_out/prof_${tag}_mppi -f \
    python tools/prof_mppi.py 65536 100 f32 6 > gpurun_out/prof_${tag}_mppi_ncu.log 2>&1
echo "mppi rc=$?"
python tools/prof_ukf.py 1048576 4 PEN_LIN > gpurun_out/prof_${tag}_ukf_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ukf_kernel -s 2 -c 2 -o gpurun_out/prof_${tag}_ukf -f \
    python tools/prof_ukf.py 1048576 4 PEN_LIN > gpurun_out/prof_${tag}_ukf_ncu.log 2>&1
echo "ukf rc=$?"
#   4. --set full of the six-state library UKF (NL6_UKF, eigen square root; the streaming kernel) at B = 2^18
#   5. --set full of the MPCB_F64_FAST rollout kernel on the config #4 shape (1024 of its 4096 controllers)
python tools/prof_ukf.py 262144 2 NL6_UKF > gpurun_out/prof_${tag}_ukf6_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ukf_stream -s 1 -c 1 -o gpurun_out/prof_${tag}_ukf6 -f \
    python tools/prof_ukf.py 262144 2 NL6_UKF > gpurun_out/prof_${tag}_ukf6_ncu.log 2>&1
echo "ukf6 rc=$?"
python tools/prof_mppi_cl.py f64fast 1024 > gpurun_out/prof_${tag}_f64fast_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mppi -s 2 -c 1 -o gpurun_out/prof_${tag}_f64fast -f \
    python tools/prof_mppi_cl.py f64fast 1024 > gpurun_out/prof_${tag}_f64fast_ncu.log 2>&1
echo "f64fast rc=$?"
