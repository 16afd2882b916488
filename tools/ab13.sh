# tiled family: rows per lane of the register tile (CARLE_TILE_R = 8: 256-row tiles, 8 warps per SM;
# 4: 128-row tiles, 16 warps per SM) x generations per temporal block (CARLE_TILE_T)
for r in 8 4; do
  for t in 16 8; do
    echo "== R=$r T=$t"
    CARLE_TILE_R=$r CARLE_TILE_T=$t python tools/biggrid.py 16384 65536 2>&1 | sed "s/^/R=$r /"
  done
done
