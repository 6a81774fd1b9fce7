for i in 1; do
  for lib in base new; do
    if [ $lib = base ]; then export YALPS_B200_LIB=$PWD/yalps_b200/libyalps_base.so; else unset YALPS_B200_LIB; fi
    echo "== $lib"
    python scripts/milp_info.py 2>&1 | cut -c1-120
    python scripts/config3_case.py SC105 16384 2>&1 | tail -2; python scripts/config3_case.py ADLITTLE 16384 2>&1 | tail -2
  done
done
