# times the partitioned lookup kernels of library variants (BPHF_LIB_VARIANT tags given as arguments; "" = the product)
for tag in "$@"; do
  echo "== variant '$tag'"; BPHF_Q_PROFILE=1 BPHF_LIB_VARIANT=$tag python tools/prof_qpart.py 1e9 134217728 2 2>&1 | grep -a "q-profile" | tail -1
done
