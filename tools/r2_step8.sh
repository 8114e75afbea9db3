for args in "2 20 0" "2 20 16" "2 20 15" "1 20 0" "1 20 16" "1 21 0" "1 21 16" "1 18 0" "1 18 15" "1 18 14" "1 18 13"; do timeout 120 python tools/msm_once.py $args 2>&1 | grep -E "curve|rror"; done
