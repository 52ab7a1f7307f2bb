NO_NCU=1 bash tests/gpu_r2_close.sh cl2 "cap64 cap96 cap128" | cut -c1-130
export DSRT_TILE_LISTS=2
NO_NCU=1 bash tests/gpu_r2_close.sh cl2a "cap48 cap96" | cut -c1-130
