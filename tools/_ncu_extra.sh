# development aid: full ncu captures of the secondary kernels (one launch each)
set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_fuse_run -s 5 -c 1 -o gpurun_out/c1_final python tools/ab_env.py c1 0 1 8 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_fuse_run -s 5 -c 1 -o gpurun_out/c3_final python tools/ab_env.py c3 0 1 8 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deform_attn_kernel -s 3 -c 1 -o gpurun_out/c4_final python tools/bench_deform.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
