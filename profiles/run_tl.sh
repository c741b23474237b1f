timeout 100 python profiles/fused_timeline.py 2>&1 | grep -v "sweep warp\|backtrack thread"
