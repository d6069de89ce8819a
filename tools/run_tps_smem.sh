# development run: per-kernel times of c4 in the reference's formats, then an ncu capture of the thread-per-stream kernels
bash tools/quick.sh "--workload c4 --n-states 2" "--workload c4 --n-states 1" "--workload c2 --n-states 2" > gpurun_out/tps_quick.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_tps -c 8 -o gpurun_out/r2_tps -f python tools/tps_profile.py > gpurun_out/tps_ncu.log 2>&1
tail -3 gpurun_out/tps_ncu.log
