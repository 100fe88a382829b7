for dt in 24 12 8; do echo "DT $dt"; LWS_K1_DT=$dt python bench.py --probes-only --probe-batch 64 2>/dev/null | grep K1 | cut -c1-160; done
