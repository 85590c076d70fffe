for wa in 64 48 32; do for tb in 32 24 16; do
  echo "WIDE_ABOVE=$wa TAIL=$tb"; GPB_WIDE_ABOVE=$wa GPB_TAIL_BLOCKS=$tb python tools/time_sizes.py 4096 8192 16384 | tail -3
done; done
