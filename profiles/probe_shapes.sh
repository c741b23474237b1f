for s in "1 824 2496" "1 824 1200" "1 800 2496" "1 832 2496" "4 824 2496" "1 416 1248" "1 824 2496 700 2000"; do MAS_B200_DEBUG=1 timeout 120 python profiles/probe_shape.py $s 2>&1 | grep -v "estimates" | tail -2; done
MAS_B200_FUSED_TEAMS=1 MAS_B200_DEBUG=1 timeout 120 python profiles/probe_shape.py 1 824 2496 2>&1 | tail -2
MAS_B200_FUSED_K=4 MAS_B200_DEBUG=1 timeout 120 python profiles/probe_shape.py 1 824 2496 2>&1 | tail -2
