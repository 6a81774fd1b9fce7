# CTA shape of the one-CTA-per-node kernel of big sparse node waves (Monster 2): YALPS_NODE_SPLIT="column warps,row groups"
for sh in 8,2 8,4 4,2 4,4 4,8 16,2 2,8 2,16; do
  echo "NODE_SPLIT $sh"; YALPS_NODE_SPLIT=$sh python scripts/milp_info.py 2>&1 | grep "Monster 2" | cut -c1-60
done
