python -m pytest tests/test_csc_gpu.py tests/test_round2_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in "" 1; do
  echo "MB200_CSC_NO_C2S=$v"
  if [ -n "$v" ]; then export MB200_CSC_NO_C2S=1; else unset MB200_CSC_NO_C2S; fi
  python profiles/scripts/time_groups.py 64 2>&1 | tail -1
  python profiles/scripts/time_csc_fused.py 200 64 2>&1 | grep "fused=True"
  python profiles/scripts/time_codes.py 2>&1 | head -1
done
