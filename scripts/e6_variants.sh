for v in "2 0" "2 1" "3 0" "3 1"; do
  set -- $v
  NLAM_NVCC_FLAGS="-DNLAM_E6_NG=$1 -DNLAM_E6_RNOW=$2" python -c "import __graft_entry__ as g; g.build(force=True)" 2>&1 | tail -2
  echo "=== NG=$1 RNOW=$2"
  python scripts/bench_inet.py g2m 32 2>&1 | grep bcast
  python scripts/bench_inet.py g2m 8 2>&1 | grep bcast
done
