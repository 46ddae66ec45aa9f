#!/bin/bash
# Round-2 profiling pass, run ON the GPU box (gpurun): launch list of the bench command, `ncu --set full` captures of the
# top kernels, and their digests.  The reports stay in /tmp (gpurun_out/ is limited to 64 MiB); what comes back are the
# text digests, the launch list and the traffic JSON bench.py reads.
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --no-c5"
OUT=gpurun_out
$CMD > $OUT/r02_plain.json 2> $OUT/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/r02_launches.csv $CMD > /tmp/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:icp_pairs_kernel -c 3 -o /tmp/r02_icp $CMD > /tmp/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:occ_ -s 10 -c 9 -o /tmp/r02_occ $CMD > /tmp/ncu_c.log 2>&1
ncu --set full --clock-control none -k regex:"normals_sweep|voxel_clouds" -c 2 -o /tmp/r02_pre $CMD > /tmp/ncu_d.log 2>&1
python profiles/c3_once.py > $OUT/r02_c3.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"icp_pairs_kernel|big_" -s 6 -c 3 -o /tmp/r02_c3 python profiles/c3_once.py > /tmp/ncu_e.log 2>&1
for k in icp occ pre c3; do python profiles/ncu_digest.py /tmp/r02_$k.ncu-rep > $OUT/r02_ncu_$k.txt 2>&1; done
python profiles/ncu_traffic.py /tmp/r02_icp.ncu-rep /tmp/r02_occ.ncu-rep /tmp/r02_c3.ncu-rep --command "$CMD" > $OUT/r02_traffic.log 2>&1
cp profiles/r02_ncu_traffic.json $OUT/ 2>/dev/null
python profiles/summarize_launches.py $OUT/r02_launches.csv > $OUT/r02_launch_list.txt 2>&1
python profiles/src_hot.py /tmp/r02_icp.ncu-rep icp_pairs_kernel 40 > $OUT/r02_src_icp.txt 2>&1
tail -3 /tmp/ncu_b.log /tmp/ncu_c.log /tmp/ncu_e.log
ls -la $OUT | tail -12
