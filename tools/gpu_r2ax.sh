# r2ax: k_shade at 10 CTAs per SM (48 registers, some spills) against the default 8 (64 registers)
bash tools/gpu_sweep.sh r2ax "OHB_X=default;OHB_BOUNCE_OCC=10;OHB_X=default;OHB_BOUNCE_OCC=10" "helmet cornell synthetic2m"
