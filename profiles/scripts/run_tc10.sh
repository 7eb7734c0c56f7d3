export MB200_SCAN_TC_STATS=1 MB200_SCAN_TC_DEBUG=0
PYTHONPATH=. timeout 600 python profiles/scripts/config5.py 100000000 2> gpurun_out/c5.err | tail -c 600
grep tcdbg gpurun_out/c5.err | awk 'NR%30==1' | head -8
grep "tensor-core" gpurun_out/c5.err | tail -2
