# Same-box A/B of two library builds (yalps_b200/libyalps_base.so = the committed build, libyalps_b200.so = the variant):
# solve() statistics of the big MILP models and one config-3 launch per Netlib base.  bash scripts/ab_bcol.sh
for i in 1; do
  for lib in base new; do
    if [ $lib = base ]; then export YALPS_B200_LIB=$PWD/yalps_b200/libyalps_base.so; else unset YALPS_B200_LIB; fi
    echo "== $lib"
    python scripts/milp_info.py 2>&1 | cut -c1-120
    python scripts/config3_case.py SC105 16384 2>&1 | tail -2; python scripts/config3_case.py ADLITTLE 16384 2>&1 | tail -2
  done
done
