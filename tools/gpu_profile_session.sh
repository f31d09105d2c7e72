#!/bin/bash
# usage: tools/gpu_profile_session.sh <tag> — plain run first, then the ncu launch list and --set full captures
tag=$1; out=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --beamset 0"
$B > $out/${tag}_plain.json 2> $out/${tag}_plain.err || { echo "plain run failed"; tail -5 $out/${tag}_plain.err; exit 1; }
$B --kernel tma > $out/${tag}_plain_tma.json 2>> $out/${tag}_plain.err || { echo "plain tma run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:b2p_fused_ldg256 -s 3 -c 2 -f -o $out/prof_${tag}_ldg256 $B > $out/${tag}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:b2p_fused_tma -s 3 -c 2 -f -o $out/prof_${tag}_tma $B --kernel tma > $out/${tag}_ncu3.log 2>&1
ls -la $out/prof_${tag}_* ; tail -2 $out/${tag}_ncu2.log
