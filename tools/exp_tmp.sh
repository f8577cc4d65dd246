tools/quick_bench.sh base | tail -1
for g in 8 16 32 64; do QLDPC_B200_OSD_GRID_B=$g tools/quick_bench.sh gridB=$g 2>&1 | tail -1; done
