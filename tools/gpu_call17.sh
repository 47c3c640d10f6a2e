mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_qmlp_sm100|k_attend_b|k_finalize_b' -c 3 -s 6 -o gpurun_out/prof_fwd_r2a -f python tools/prof_bags.py > gpurun_out/ncu_r2a.log 2>&1; tail -3 gpurun_out/ncu_r2a.log
DSMIL_B200_PAIR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_fwd_pair' -c 1 -s 2 -o gpurun_out/prof_pair_r2a -f python tools/prof_bags.py > gpurun_out/ncu_r2b.log 2>&1; tail -3 gpurun_out/ncu_r2b.log
ls -la gpurun_out/*.ncu-rep | tail -3
