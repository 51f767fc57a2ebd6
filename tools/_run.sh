mkdir -p gpurun_out/s18
PCLS_LIB_SUFFIX=_dbg timeout 300 python tools/tc_debug_run.py darknet21_kitti_64x2048_b32 > gpurun_out/s18/dbg_dk21.txt 2>&1
PCLS_LIB_SUFFIX=_dbg timeout 300 python tools/tc_debug_run.py squeezesegv2_kitti_64x2048_b32 > gpurun_out/s18/dbg_ssv2.txt 2>&1
tail -3 gpurun_out/s18/dbg_dk21.txt
