#!/bin/bash
# usage: tools/gpu_round_session.sh <tag> — one 1-GPU box: GPU tests, smoke, the two bench arms as the
# driver runs them, then the ncu launch list and the --set full captures (each after its plain run)
tag=$1; out=gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,pci.bus_id,clocks.max.sm --format=csv; } > $out/${tag}_box.txt 2>&1
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "gpu tests rc=$? $((SECONDS-t0)) s"; tail -3 $out/${tag}_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
t0=$SECONDS
timeout 600 python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference arm rc=$? $((SECONDS-t0)) s"
t0=$SECONDS
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$? $((SECONDS-t0)) s"; tail -3 $out/${tag}_bench_n1.err
head -c 1500 $out/${tag}_bench_n1.json; echo
timeout 900 bash tools/gpu_profile_session.sh $tag
