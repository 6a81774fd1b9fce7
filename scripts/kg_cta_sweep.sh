export YALPS_CASE_PATH=7
for c in 148 147 146 140 134 129 128 120 112; do
  echo "CTAS $c"; YALPS_KG_CTAS=$c python scripts/k4_case.py 1024 2048 200 | tail -1
  YALPS_KG_CTAS=$c python scripts/k4_case.py 0 0 8192 25FV47 | tail -1
done
