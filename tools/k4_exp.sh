for cfg in "DH_LOSS_GROUPS=2" "DH_LOSS_GROUPS=2 DH_LOSS_MEM=1" "DH_LOSS_GROUPS=1" "DH_LOSS_GROUPS=1 DH_LOSS_MEM=1"; do
  echo "== $cfg"
  env $cfg python tools/k4_scaling.py 2>&1 | grep -E "resize|all three"
done
