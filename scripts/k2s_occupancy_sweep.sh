# CTAs per SM of the config-3 kernel (K2s, 64 threads): python scripts/config3_case.py under YALPS_CTAS_PER_SM
for c in 16 12 10 8 7 6 5 4 3; do
  echo "CTAS_PER_SM $c"
  YALPS_CTAS_PER_SM=$c python scripts/config3_case.py SC105 16384 | tail -n 2 | head -1
  YALPS_CTAS_PER_SM=$c python scripts/config3_case.py ADLITTLE 16384 | tail -n 2 | head -1
done
