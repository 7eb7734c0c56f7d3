# many-group CSC path: parity tests, step time at 64 groups (100 and 200 bp), fp32 code retrieval, ncu launch list at 200 bp
python -m pytest tests/test_csc_gpu.py tests/test_round2_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
python profiles/scripts/time_groups.py 64 2>&1 | tail -1
python profiles/scripts/time_csc_fused.py 200 64 2>&1 | grep "fused=True"
python profiles/scripts/time_codes.py 2>&1 | head -1
bash profiles/scripts/r02_groups_launches.sh 200 64 2>/dev/null | head -${1:-14}
