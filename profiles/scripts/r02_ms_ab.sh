# k_mask_scale_g A/B on one box: in-tree library vs build_prev/lib_prev.so (another MG_MINB); code retrieval rate, 64-group step, kernel time under ncu
for lib in "" build_prev/lib_prev.so; do
  export MB200_LIBRARY=$lib; echo "library: ${lib:-in-tree}"
  python profiles/scripts/time_codes.py 2>&1 | head -1
  python profiles/scripts/time_groups.py 64 2>&1 | tail -1
  python profiles/scripts/time_csc_fused.py 200 64 2>&1 | grep "fused=True"
  bash profiles/scripts/codes_kernel_times.sh 2>/dev/null | grep -E "k_mask_scale_g|launches"
done
unset MB200_LIBRARY
python -m pytest tests/test_csc_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
