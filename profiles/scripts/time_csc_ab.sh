# A/B of the fused CSC step on ONE box: the in-tree library against other builds of the same ABI
for rep in 1 2; do
  for lib in "" "$@"; do
    MB200_LIBRARY=$lib timeout 300 python profiles/scripts/time_csc_fused.py 100 2>&1 | grep "fused=True" | sed "s|^|${lib:-in-tree} |"
  done
done
