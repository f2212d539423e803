#!/bin/bash
# Evidence run on the GPU box (under gpurun, from the repo root): smoke, the bench line and its reference arm, one
# `ncu --set full` capture per bench configuration (raw + source pages exported to CSV here, so that the reports need
# not travel), the SASS of the headline kernel and the ncu launch list of the bench command.
#   gpurun --timeout 2400 -- 'bash tools/evidence_run.sh r02p'
out=gpurun_out/${1:-evidence}
mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
python bench.py > $out/bench_1gpu.json 2> $out/bench_1gpu.err
python bench.py --impl reference > $out/bench_reference.json 2>> $out/bench_1gpu.err
cd $out
P="python ../../tools/profile_run.py"
N="ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:simb_step -f"
cap() {
    tag=$1; shift
    $N -c $1 -o $tag $P "${@:2}" > $tag.log 2>&1
    ncu -i $tag.ncu-rep --page raw --csv > ${tag}_raw.csv 2>/dev/null
    ncu -i $tag.ncu-rep --page source --csv --print-source cuda,sass > ${tag}_source.csv 2>/dev/null
    rm -f $tag.ncu-rep
}
cap fused 2 --network mm1 --replications 1048576 --k 128 --warm-until 300 --launches 1
cap k1 4 --network mm1 --replications 1048576 --k 1 --warm-until 300 --launches 2
cap k1_2p22 4 --network mm1 --replications 4194304 --k 1 --warm-until 300 --launches 2
cap cfg_tandem 2 --network tandem --replications 4194304 --k 128 --warm-until 50 --launches 1
cap cfg_gateway_network 2 --network gateway_network --replications 1048576 --k 128 --warm-until 50 --launches 1
cap cfg_large_mixed 2 --network large_mixed --replications 2097152 --k 128 --warm-until 3 --launches 1
cuobjdump -sass -fun "_ZN4simb16simb_step_kernelILi0ELb0ELi0ELb0ELb1ELb0EEEv11SimbNetDesc12SimbDevState10SimbLaunchj" ../../sim_b200/libsimb200.so > fused_sass.txt 2>&1
cd ../..
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --strong-total-log2 23 > $out/bench_under_ncu.log 2>&1
gzip -f $out/*_source.csv
ls -la $out
du -sh gpurun_out
cat $out/smoke.log
