# 64-thread K2s built for 8 / 10 / 12 CTAs per SM (register caps 128 / 96 / 80): config 3
for lib in base v10 v12; do
  export YALPS_B200_LIB=$PWD/yalps_b200/libyalps_$lib.so
  echo "== $lib"
  python scripts/config3_case.py SC105 16384 | tail -n 2 | head -1
  python scripts/config3_case.py ADLITTLE 16384 | tail -n 2 | head -1
  python scripts/config3_case.py SC105 32768 | tail -n 2 | head -1
done
